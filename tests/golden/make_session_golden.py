"""Generates tests/golden/session_golden.npz from the REFERENCE ITSELF (run in the build container).

The reference's session layer (Inference/PythonInference/offline_asr_session.py, stream_asr_session.py, vad/src/vad.py,
asr/src/asr.py, punc_recover/src/punc_recover.py) is imported UNMODIFIED from /root/reference and run on a test recording; the only
substitutions are the three third-party modules this container lacks: `onnxruntime` (-> the reference's own vendored onnxruntime
1.10.0 binary through oracle/ort_ref.py), `librosa` (-> wave reader, the recording is 16 kHz already) and `soundfile` (unused).

Part A (real models).  The recording: faint noise / the reference wav (4.2 s) / noise / the first 2.6 s of the wav, twice over with
different gaps (about 20 s, four sentences): begin / end events of the streaming session and the > 15 s split of the offline
segmenter (offline_asr_session.py:184-217) are exercised.

Part B (state machines only).  The reference's ASRSession / TaskContent / OfflineVAD driven by SCRIPTED voice-activity patterns (a
stub VAD that returns the scripted 0/1 decisions, a stub ASR that returns a description of what it was given): event traces that
cover the paths real audio rarely reaches (inter-break, result-change accumulation, chunk resets, final_send).
"""
import json
import os
import sys
import types
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ort_ref  # noqa: E402

REF = "/root/reference/Inference/PythonInference"


class _IO:
    def __init__(self, name):
        self.name = name


class _Session:
    """onnxruntime.InferenceSession look-alike over the vendored ORT binary."""
    _names = {"encoder.onnx": (["inputs"], ["Identity:0"]), "ctc_model.onnx": (["inputs"], ["Identity:0"]),
              "translator.onnx": (["inputs", "enc"], ["Identity:0"]), "vad.onnx": (["inputs"], ["output_0"]),
              "punc.onnx": (["inputs", "mask", "encoder/strided_slice_1/input:0"], ["Identity:0"])}

    def __init__(self, path, *a, **k):
        self.m = ort_ref.OrtModel(os.path.join(REF, path) if not os.path.isabs(path) else path, 1)
        self.ins, self.outs = self._names[os.path.basename(path)]

    def get_inputs(self):
        return [_IO(n) for n in self.ins]

    def get_outputs(self):
        return [_IO(n) for n in self.outs]

    def run(self, out_names, input_feed):
        return [self.m.run(dict(input_feed), out_names[0])]


def _read_wav(path):
    w = wave.open(path)
    return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768


def _install_stubs():
    ort = types.ModuleType("onnxruntime")
    ort.InferenceSession = _Session
    sys.modules["onnxruntime"] = ort
    lib = types.ModuleType("librosa")
    lib.load = lambda path, sr=16000, **k: (_read_wav(path), sr)
    sys.modules["librosa"] = lib
    sys.modules["soundfile"] = types.ModuleType("soundfile")


def build_recording():
    x = _read_wav(os.path.join(ROOT, "tests/golden/BAC009S0764W0121.wav"))
    rng = np.random.default_rng(11)
    sil = lambda s: (rng.standard_normal(int(s * 16000)) * 2e-4).astype(np.float32)
    rec = np.concatenate([sil(0.7), x, sil(1.3), x[:int(2.6 * 16000)], sil(1.0), x, sil(2.5), x[:int(2.6 * 16000)], sil(1.0)])
    rec = rec[:len(rec) // 160 * 160]        # (the reference's own reshape at offline_asr_session.py:82 needs a whole number of 160-sample frames)
    return np.clip(np.round(rec * 32768), -32768, 32767).astype("<i2")


class ScriptVAD:
    """Stub for vad/src/vad.py: the k-th call answers with the k-th scripted block of ten 0/1 decisions (as logits +-1) in the LAST ten
    frames (stream_asr_session.py:341 keeps only those); earlier frames are -1."""

    def __init__(self, script):
        self.script, self.k = script, 0

    def inference(self, wav):
        n = wav.shape[1]
        o = -np.ones((1, n, 1), np.float32)
        blk = self.script[min(self.k, len(self.script) - 1)]
        o[0, n - 10:, 0] = np.where(np.asarray(blk) > 0, 1.0, -1.0)
        self.k += 1
        return o


class ScriptASR:
    """Stub for asr/src/asr.py: the 'encoder output' of a chunk is its length, the 'text' lists what decode() was handed."""

    def extract_feature(self, wav):
        return np.asarray([[len(wav)]], np.int64)

    def decode(self, feats):
        return "decode(" + ",".join(str(int(f[0, 0])) for f in feats) + ")"


class ScriptPunc:
    def punc_recover(self, t):
        return list(t) + ["<p>"]


def make_scripts():
    """Seeded voice-activity scripts: one block of ten decisions per 100 ms VAD call."""
    rng = np.random.default_rng(5)
    scripts = []
    for n in range(6):
        blocks, state = [], 0
        for _ in range(70 + 10 * n):
            if rng.random() < (0.12 if state else 0.2):
                state ^= 1
            p = 0.93 if state else 0.06
            if rng.random() < 0.25:                        # ragged blocks around the thresholds (5 / 8 of 10)
                p = rng.choice([0.3, 0.5, 0.7])
            blocks.append((rng.random(10) < p).astype(np.int32))
        scripts.append(np.stack(blocks))
    return scripts


def script_pcm(n):
    """Audio of the scripted runs: its content is irrelevant (the VAD is scripted), only its length matters -- a formula, not stored."""
    return ((np.arange(n, dtype=np.int64) * 7919) % 6001 - 3000).astype("<i2")


def scripted(off, stream):
    out = {}
    for si, script in enumerate(make_scripts()):
        # streaming session: 20 ms packets (320 samples @ 16 kHz), one VAD call per 100 ms
        ss = stream.ASRSession()
        ss.asr, ss.punc = ScriptASR(), ScriptPunc()
        ss.task_content.compile(ScriptVAD(script))
        pcm = script_pcm(len(script) * 1600)
        events = []
        for k, p in enumerate(range(0, len(pcm), 320)):
            r = ss.send(pcm[p:p + 320].tobytes())
            if r is not None:
                events.append({"packet": k, **r})
        r = ss.final_send()
        if r is not None:
            events.append({"packet": -1, **r})
        out[f"script{si}"] = script.astype(np.int8)
        out[f"script{si}_events"] = np.frombuffer(json.dumps(events, ensure_ascii=False).encode("utf-8"), dtype=np.uint8)
        print("script", si, len(events), "events", sorted({e["event_type"] for e in events}))
        # offline segmenter on the same decisions (one decision per 10 ms frame)
        ov = off.OfflineVAD(sr=16000)

        class _Whole:
            def inference(self, wav, s=script):
                n = wav.shape[1]
                d = np.concatenate([s.reshape(-1), np.zeros(max(n - s.size, 0), np.int32)])[:n]
                return np.where(d > 0, 1.0, -1.0).astype(np.float32).reshape(1, n, 1)
        ov.compile(_Whole())
        wav = (pcm.astype(np.float32) / 32768)[:len(pcm) // 80 * 80]
        out[f"script{si}_offline"] = np.asarray(ov.vad(wav), dtype=np.float64).reshape(-1, 2)
    # the offline merge / split rule on hand-made segment lists (offline_asr_session.py:184-217)
    ov = off.OfflineVAD(sr=16000)
    cases = [[[0.5, 3.0], [3.05, 6.0], [6.3, 9.0]], [[0.0, 16.0], [16.05, 31.0]], [[1.0, 47.5], [48.0, 49.0]], [[0.2, 30.2], [31.0, 32.0]],
             [[0.0, 2.0], [2.05, 14.0], [14.08, 20.0], [20.5, 21.0]]]
    out["recover_cases"] = np.frombuffer(json.dumps([[c, ov.recover([list(x) for x in c])] for c in cases]).encode(), dtype=np.uint8)
    return out


PUNC_TEXTS = ["甚至出现交易几乎停止的情况", "甚至出现交易几乎停制的情况甚至出现交易几乎品甚至出现交易几乎停制的情况甚至出现交易几乎挺",
              "今天天气怎么样", "你好请问你叫什么名字我想知道明天会不会下雨如果下雨的话我们就不去公园了", "我",
              "他说这个问题很难解决但是我们必须想办法因为时间已经不多了你觉得呢"]


def punctuation(stream):
    """punc_recover/src/punc_recover.py on a few sentences: token ids, the model's class probabilities, the punctuated token list."""
    import importlib
    pr = importlib.import_module("punc_recover.src.punc_recover")
    from utils.user_config import UserConfig
    cfg = UserConfig("./punc_recover/src/configs/data.yml", "./punc_recover/src/configs/punc_settings.yml")
    punc = pr.Punc(cfg)
    out, cases = {}, []
    for i, txt in enumerate(PUNC_TEXTS):
        x = [punc.vocab_featurizer.startid()] + punc.vocab_featurizer.extract(txt) + [punc.vocab_featurizer.endid()]
        x = np.array([x], "int32")
        names = [n.name for n in punc.model.get_inputs()]
        probs = punc.model.run([punc.model.get_outputs()[0].name],
                               input_feed={names[0]: x, names[1]: punc.creat_mask(x), names[2]: punc.pos_encode_inputs})[0]
        out[f"punc{i}_ids"] = x[0]
        out[f"punc{i}_probs"] = probs[0].astype(np.float32)
        cases.append([txt, punc.punc_recover(txt)])
        print("punc", "".join(cases[-1][1]))
    out["punc_cases"] = np.frombuffer(json.dumps(cases, ensure_ascii=False).encode("utf-8"), dtype=np.uint8)
    out["punc_pe"] = punc.pos_encode_inputs[0, :64].astype(np.float32)          # first rows of the positional table the caller feeds
    return out


def main():
    _install_stubs()
    os.chdir(REF)
    sys.path.insert(0, REF)
    import importlib
    off = importlib.import_module("offline_asr_session")
    stream = importlib.import_module("stream_asr_session")

    pcm = build_recording()
    tmp = "/tmp/session_golden.wav"
    with wave.open(tmp, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(pcm.tobytes())
    out = {"pcm": pcm}

    # ---- VAD model alone (vad/src/vad.py:24-28): the offline session's call on the whole recording
    wav = pcm.astype(np.float32) / 32768
    wav = wav[:len(wav) // 80 * 80]
    vad_sess = _Session("./vad/models/vad.onnx")
    frames = wav[::2].reshape(1, -1, 80)
    out["vad_logits"] = vad_sess.run(["output_0"], {"inputs": frames.astype(np.float32)})[0].reshape(-1).astype(np.float32)

    # ---- offline session (offline_asr_session.py:37-50)
    s = off.ASRSession()
    out["offline_segments"] = np.asarray(s.offline_vad.vad(wav), dtype=np.float64)
    responses = s.send(tmp)
    out["offline_responses"] = np.frombuffer(json.dumps(responses, ensure_ascii=False).encode("utf-8"), dtype=np.uint8)
    # the same segments without punctuation (the phone -> character text the ASR path alone produces)
    plain = []
    for (b, e) in out["offline_segments"]:
        data = wav[int(b * 16000):int(e * 16000)]
        plain.append(s.asr.decode([s.asr.extract_feature(data)]))
    out["offline_plain_text"] = np.frombuffer(json.dumps(plain, ensure_ascii=False).encode("utf-8"), dtype=np.uint8)

    # ---- streaming session (stream_asr_session.py:106-262), 20 ms packets of the generator at :462-478 (160 samples each)
    class _NoPunc:
        def punc_recover(self, t):
            return t
    for tag, punc in (("stream", None), ("stream_nopunc", _NoPunc())):
        ss = stream.ASRSession()
        if punc is not None:
            ss.punc = punc
        events = []
        for p in range(0, len(pcm), 160):
            r = ss.send(pcm[p:p + 160].tobytes())
            if r is not None:
                events.append({"packet": p // 160, **r})
        r = ss.final_send()
        if r is not None:
            events.append({"packet": -1, **r})
        out[tag + "_events"] = np.frombuffer(json.dumps(events, ensure_ascii=False).encode("utf-8"), dtype=np.uint8)
        print(tag, json.dumps(events, ensure_ascii=False, indent=1))
    out.update(scripted(off, stream))
    out.update(punctuation(stream))
    np.savez_compressed(os.path.join(ROOT, "tests/golden/session_golden.npz"), **out)
    print("segments", out["offline_segments"].tolist())
    print("responses", json.dumps(responses, ensure_ascii=False))
    print("plain", plain)


if __name__ == "__main__":
    main()
