"""Generates tests/golden/session_golden.npz from the REFERENCE ITSELF (run in the build container).

The reference's session layer (Inference/PythonInference/offline_asr_session.py, stream_asr_session.py, vad/src/vad.py,
asr/src/asr.py, punc_recover/src/punc_recover.py) is imported UNMODIFIED from /root/reference and run on a test recording; the only
substitutions are the three third-party modules this container lacks: `onnxruntime` (-> the reference's own vendored onnxruntime
1.10.0 binary through oracle/ort_ref.py), `librosa` (-> wave reader, the recording is 16 kHz already) and `soundfile` (unused).

The recording: 0.7 s of faint noise, the reference wav (4.2 s), 1.3 s of faint noise, the first 2.6 s of the wav again, 1.0 s of
faint noise -- two sentences, so that begin / change / inter-break / end events and the offline segment merge are all exercised.
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
    rec = np.concatenate([sil(0.7), x, sil(1.3), x[:int(2.6 * 16000)], sil(1.0)])
    return np.clip(np.round(rec * 32768), -32768, 32767).astype("<i2")


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
    np.savez_compressed(os.path.join(ROOT, "tests/golden/session_golden.npz"), **out)
    print("segments", out["offline_segments"].tolist())
    print("responses", json.dumps(responses, ensure_ascii=False))
    print("plain", plain)


if __name__ == "__main__":
    main()
