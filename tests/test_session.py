"""Session layer on the CPU: the voice-activity state machines of tensorflowasr_b200/session.py against event traces
produced by the REFERENCE's own session classes (tests/golden/make_session_golden.py part B: scripted VAD decisions, stub recogniser),
and the VAD oracle against the reference's onnxruntime output.  The GPU side (real VAD + recogniser) is tests/test_gpu_session.py."""
import json
import os

import numpy as np
import pytest

from tensorflowasr_b200 import session as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests/golden/session_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _json(a):
    return json.loads(bytes(a).decode("utf-8"))


def script_pcm(n):
    return ((np.arange(n, dtype=np.int64) * 7919) % 6001 - 3000).astype("<i2")


class ScriptVAD:
    """k-th call -> the k-th scripted block of ten decisions in the last ten frames (same stub as the golden generator's)."""

    def __init__(self, script):
        self.script, self.k = script, 0

    def inference(self, wav):
        n = wav.shape[1]
        o = -np.ones((1, n, 1), np.float32)
        o[0, n - 10:, 0] = np.where(np.asarray(self.script[min(self.k, len(self.script) - 1)]) > 0, 1.0, -1.0)
        self.k += 1
        return o


class ScriptASR:
    def extract_feature(self, wav):
        return np.asarray([[len(wav)]], np.int64)

    def decode(self, feats):
        return "decode(" + ",".join(str(int(f[0, 0])) for f in feats) + ")"


class ScriptPunc:
    def punc_recover(self, t):
        return list(t) + ["<p>"]


@pytest.mark.parametrize("si", range(6))
def test_stream_session_events_match_reference(gold, si):
    script = gold[f"script{si}"].astype(np.int32)
    want = _json(gold[f"script{si}_events"])
    sess = S.StreamASRSession(ScriptASR(), ScriptVAD(script), ScriptPunc())
    pcm = script_pcm(len(script) * 1600)
    got = []
    for k, p in enumerate(range(0, len(pcm), 320)):
        r = sess.send(pcm[p:p + 320].tobytes())
        if r is not None:
            got.append({"packet": k, **r})
    r = sess.final_send()
    if r is not None:
        got.append({"packet": -1, **r})
    assert got == want
    assert {e["event_type"] for e in got} >= {"sentence begin", "sentence end"}


def test_inter_break_is_covered(gold):
    kinds = set()
    for si in range(6):
        kinds |= {e["event_type"] for e in _json(gold[f"script{si}_events"])}
    assert "inter break" in kinds


@pytest.mark.parametrize("si", range(6))
def test_offline_segmenter_matches_reference(gold, si):
    script = gold[f"script{si}"].astype(np.int32)

    class Whole:
        def inference(self, wav):
            n = wav.shape[1]
            d = np.concatenate([script.reshape(-1), np.zeros(max(n - script.size, 0), np.int32)])[:n]
            return np.where(d > 0, 1.0, -1.0).astype(np.float32).reshape(1, n, 1)
    ov = S.OfflineVAD(sr=16000)
    ov.compile(Whole())
    wav = script_pcm(len(script) * 1600).astype(np.float32) / 32768
    got = np.asarray(ov.vad(wav), dtype=np.float64).reshape(-1, 2)
    np.testing.assert_array_equal(got, gold[f"script{si}_offline"])


def test_recover_merge_and_split(gold):
    ov = S.OfflineVAD(sr=16000)
    for segs, want in _json(gold["recover_cases"]):
        assert ov.recover([list(x) for x in segs]) == want


def test_offline_session_with_stub_models(gold):
    """send(): trims to whole frames, one response per segment with millisecond bounds (offline_asr_session.py:37-50)."""
    script = gold["script0"].astype(np.int32)

    class Whole:
        def inference(self, wav):
            n = wav.shape[1]
            d = np.concatenate([script.reshape(-1), np.zeros(max(n - script.size, 0), np.int32)])[:n]
            return np.where(d > 0, 1.0, -1.0).astype(np.float32).reshape(1, n, 1)
    sess = S.OfflineASRSession(ScriptASR(), Whole(), ScriptPunc())
    wav = script_pcm(len(script) * 1600 + 37).astype(np.float32) / 32768
    resp = sess.send(wav)
    (s, e), = gold["script0_offline"]
    assert len(resp) == 1 and resp[0]["sentence_begin_time"] == int(s * 1000) and resp[0]["sentence_end_time"] == int(e * 1000)
    n = len(wav) // 160 * 160
    assert resp[0]["best_text"][-1] == "<p>" and "".join(resp[0]["best_text"][:-1]) == f"decode({len(wav[:n][int(s * 16000):int(e * 16000)])})"


def test_vad_oracle_is_pinned_to_the_reference(gold):
    """oracle/vad_ref.py against the reference's onnxruntime on vad.onnx (golden 'vad_logits')."""
    from oracle import ort_ref, vad_ref
    from tensorflowasr_b200 import vad_model as V
    path = os.path.join(ort_ref.REF_DIR, "models", "vad", "vad.onnx")
    if not os.path.isfile(path):
        pytest.skip("oracle/_ref/models/vad/vad.onnx not staged (run oracle/build_ref.py)")
    raw = V.import_vad(path)
    pcm = gold["pcm"].astype(np.float32) / 32768
    y = vad_ref.vad_forward(raw, pcm[::2].reshape(-1, 80))
    assert np.abs(y - gold["vad_logits"]).max() < 5e-5
    dev = V.vad_device_tensors(raw)
    assert dev["c0.w"].shape == (80, 400) and dev["d4.w"].shape == (4, 80) and np.all(dev["d4.w"][1:] == 0)
    np.testing.assert_array_equal(dev["c0.w"][:, 80:160], raw["c0.w"][:, :, 1])


def test_punctuation_oracle_is_pinned_to_the_reference(gold):
    """oracle/punc_ref.py against the reference's Punc class on its own onnxruntime (golden 'punc*_probs'), and the importer's layout."""
    from oracle import ort_ref, punc_ref
    from tensorflowasr_b200 import punc_model as P
    path = os.path.join(ort_ref.REF_DIR, "models", "punc", "punc.onnx")
    if not os.path.isfile(path):
        pytest.skip("oracle/_ref/models/punc/punc.onnx not staged (run oracle/build_ref.py)")
    raw = P.import_punc(path)
    pe = P.punc_positional_encoding()
    np.testing.assert_array_equal(pe[:64], gold["punc_pe"])
    i = 0
    while f"punc{i}_ids" in gold:
        y = punc_ref.punc_forward(raw, gold[f"punc{i}_ids"], pe)
        assert np.abs(y - gold[f"punc{i}_probs"]).max() < 1e-5
        i += 1
    assert i >= 5
    dev = P.punc_device_tensors(raw)
    assert dev["l0.qkv.w"].shape == (192, 64) and dev["c1.w"].shape == (64, 192) and dev["up.w"].shape == (768, 64)
    np.testing.assert_allclose(dev["l2.qkv.w"][:64], raw["l2.q.w"].T / np.sqrt(8.0), rtol=1e-6)
    np.testing.assert_array_equal(dev["c0.w"][:, 64:128], raw["c0.w"][:, :, 1])


def test_stream_session_tolerates_ragged_packets(gold):
    """Packets that are not whole 160-sample frames (100 samples each): the reference's reshape would raise; here the VAD window drops
    the ragged head and the session keeps producing well-formed events."""
    script = gold["script2"].astype(np.int32)
    sess = S.StreamASRSession(ScriptASR(), ScriptVAD(script), None)
    pcm = script_pcm(len(script) * 1600)
    events = []
    for p in range(0, len(pcm), 100):
        r = sess.send(pcm[p:p + 100].tobytes())
        if r is not None:
            events.append(r)
    r = sess.final_send()
    if r is not None:
        events.append(r)
    kinds = [e["event_type"] for e in events]
    assert kinds.count("sentence begin") >= 1 and kinds.count("sentence end") >= 1
    for e in events:
        if e["event_type"] == "sentence end":
            assert e["sentence_end_time"] >= e["sentence_begin_time"] and e["best_text"].startswith("decode(")
    assert sess.send(b"") is None          # an empty packet is a no-op


def test_offline_vad_without_speech_returns_nothing():
    class Quiet:
        def inference(self, wav):
            return -np.ones((1, wav.shape[1], 1), np.float32)
    ov = S.OfflineVAD(sr=16000)
    ov.compile(Quiet())
    assert ov.vad(np.zeros(32000, np.float32)) == []
    sess = S.OfflineASRSession(ScriptASR(), Quiet(), None)
    assert sess.send(np.zeros(32000 + 77, np.float32)) == []
