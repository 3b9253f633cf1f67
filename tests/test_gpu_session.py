"""Session layer on the GPU: the voice-activity model through the C ABI (b200asr_vad_*) against the reference's
onnxruntime output, and the offline / streaming sessions end to end (VAD gating + recogniser + translator) against the events the
REFERENCE's own session classes produced on the same recording (tests/golden/make_session_golden.py part A)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests/golden/session_golden.npz"))


@pytest.fixture(scope="module")
def staged():
    from oracle import ort_ref
    need = [os.path.join(ort_ref.REF_DIR, "models", "vad", "vad.onnx"), os.path.join(ort_ref.REF_DIR, "dict", "pinyin.txt"),
            os.path.join(ort_ref.REF_DIR, "dict", "lm_tokens.txt")]
    if not all(os.path.isfile(p) for p in need) or ort_ref.model_dir("streaming") is None:
        pytest.skip("reference models / vocabularies not staged (oracle/build_ref.py)")
    return ort_ref


@pytest.fixture(scope="module")
def vad(staged):
    from tensorflowasr_b200 import vad_model as V
    return V.VAD(model_path=os.path.join(staged.REF_DIR, "models", "vad", "vad.onnx"))


def _json(a):
    return json.loads(bytes(a).decode("utf-8"))


def _asr(staged, kind):
    from tensorflowasr_b200 import asr as A
    cfg = {"running_config": {}, "optimizer_config": {}, "model_config": {},
           "speech_config": {"sample_rate": 16000, "frame_ms": 25, "stride_ms": 10, "num_feature_bins": 80, "streaming": True, "streaming_bucket": 0.5},
           "inp_config": {"vocabulary": os.path.join(staged.REF_DIR, "dict", "pinyin.txt"), "blank_at_zero": False, "beam_width": 1},
           "tar_config": {"vocabulary": os.path.join(staged.REF_DIR, "dict", "lm_tokens.txt"), "blank_at_zero": False, "beam_width": 1}}
    a = A.ASR(cfg)
    a.compile(staged.model_dir(kind), chunked=False)
    return a


def test_vad_logits_match_reference(vad, gold):
    """b200asr_vad_infer on the whole recording (the offline session's call) vs onnxruntime on vad.onnx: exact-fp32 path, 1e-4."""
    pcm = gold["pcm"].astype(np.float32) / 32768
    frames = pcm[::2].reshape(1, -1, 80)
    got = vad.inference(frames)
    assert got.shape == (1, frames.shape[1], 1)
    want = gold["vad_logits"]
    assert np.abs(got.reshape(-1) - want).max() < 1e-4
    np.testing.assert_array_equal(got.reshape(-1) >= 0, want >= 0)


def test_vad_batch_stride_and_oracle(vad, gold, staged):
    """Several sessions in one call, 16 kHz input with stride 2, ragged sizes; every row equals the row alone and the oracle."""
    import torch
    from oracle import vad_ref
    from tensorflowasr_b200 import vad_model as V
    raw = V.import_vad(os.path.join(staged.REF_DIR, "models", "vad", "vad.onnx"))
    pcm = gold["pcm"].astype(np.float32) / 32768
    rng = np.random.default_rng(3)
    for B, N in ((1, 1), (3, 7), (5, 300), (2, 1001)):
        starts = rng.integers(0, len(pcm) - N * 160, size=B)
        wav16 = np.stack([pcm[s:s + N * 160] for s in starts])
        x = torch.from_numpy(wav16).cuda()
        got = vad.model.infer(x, stride=2).cpu().numpy()
        for b in range(B):
            want = vad_ref.vad_forward(raw, wav16[b, ::2].reshape(-1, 80))
            assert np.abs(got[b] - want).max() < 1e-4, (B, N, b)
        alone = vad.model.infer(x[:1].contiguous(), stride=2).cpu().numpy()
        np.testing.assert_array_equal(alone[0], got[0])
        got8 = vad.model.infer(torch.from_numpy(np.ascontiguousarray(wav16[:, ::2])).cuda(), stride=1).cpu().numpy()
        np.testing.assert_array_equal(got8, got)
    with pytest.raises(ValueError):
        vad.model.infer(torch.zeros((1, 100), device="cuda"))
    lib, h = vad.model.lib, vad.model._h
    assert lib.b200asr_vad_infer(h, x.data_ptr(), 1, 1, 3, x.data_ptr(), None) != 0 and b"stride" in lib.b200asr_last_error(h)
    assert lib.b200asr_recognize(h, x.data_ptr(), 1, 1600, x.data_ptr(), x.data_ptr(), None) != 0          # a VAD handle is not a recogniser


def test_stream_session_matches_reference_events(vad, gold, staged):
    """20 ms... the reference generator's 160-sample packets through StreamASRSession == the events of the reference's ASRSession
    (same packets, punctuation off): begin / inter-break / end, millisecond bounds and characters."""
    from tensorflowasr_b200 import session as S
    sess = S.StreamASRSession(_asr(staged, "streaming"), vad)
    pcm = gold["pcm"]
    got = []
    for p in range(0, len(pcm), 160):
        r = sess.send(pcm[p:p + 160].tobytes())
        if r is not None:
            got.append({"packet": p // 160, **r})
    r = sess.final_send()
    if r is not None:
        got.append({"packet": -1, **r})
    want = _json(gold["stream_nopunc_events"])
    assert [(e["packet"], e["event_type"]) for e in got] == [(e["packet"], e["event_type"]) for e in want]
    assert got == want


def test_offline_session_matches_reference(vad, gold, staged):
    from tensorflowasr_b200 import session as S
    sess = S.OfflineASRSession(_asr(staged, "offline"), vad)
    wav = gold["pcm"].astype(np.float32) / 32768
    resp = sess.send(wav)
    segs = gold["offline_segments"]
    assert [[r["sentence_begin_time"], r["sentence_end_time"]] for r in resp] == [[int(s * 1000), int(e * 1000)] for s, e in segs]
    assert [r["best_text"] for r in resp] == _json(gold["offline_plain_text"])


@pytest.fixture(scope="module")
def punc(staged):
    from tensorflowasr_b200 import punc_model as P
    need = [os.path.join(staged.REF_DIR, "models", "punc", "punc.onnx"), os.path.join(staged.REF_DIR, "dict", "lm_tokens_ch.txt"),
            os.path.join(staged.REF_DIR, "dict", "lm_tokens_bd.txt")]
    if not all(os.path.isfile(p) for p in need):
        pytest.skip("punctuation model / vocabularies not staged")
    cfg = {"running_config": {}, "model_config": {"d_model": 64, "pe_input": 1024},
           "punc_vocab": {"vocabulary": need[1], "blank_at_zero": True, "beam_width": 1},
           "punc_biaodian": {"vocabulary": need[2], "blank_at_zero": True, "beam_width": 1}}
    return P.Punc(cfg, model_path=need[0])


def test_punctuation_matches_reference(punc, gold):
    """b200asr_punc_infer vs the reference's Punc class on its onnxruntime: probabilities (1e-5) and the punctuated sentences."""
    import torch
    i = 0
    while f"punc{i}_ids" in gold:
        ids = torch.from_numpy(gold[f"punc{i}_ids"].astype(np.int32)).cuda()
        got = punc.model.infer(ids).cpu().numpy()
        assert np.abs(got - gold[f"punc{i}_probs"]).max() < 1e-5, i
        i += 1
    for txt, want in _json(gold["punc_cases"]):
        assert punc.punc_recover(txt) == want
    with pytest.raises(RuntimeError):
        punc.model.infer(torch.tensor([1, 999999, 2], dtype=torch.int32, device="cuda"))
    got = punc.model.infer(torch.from_numpy(gold["punc0_ids"].astype(np.int32)).cuda()).cpu().numpy()      # the handle survives the error
    assert np.abs(got - gold["punc0_probs"]).max() < 1e-5


def test_sessions_with_punctuation_match_reference(vad, punc, gold, staged):
    """The complete sessions (VAD + recogniser + translator + punctuation), nothing on onnxruntime: the reference's own events."""
    from tensorflowasr_b200 import session as S
    sess = S.StreamASRSession(_asr(staged, "streaming"), vad, punc)
    pcm = gold["pcm"]
    got = []
    for p in range(0, len(pcm), 160):
        r = sess.send(pcm[p:p + 160].tobytes())
        if r is not None:
            got.append({"packet": p // 160, **r})
    r = sess.final_send()
    if r is not None:
        got.append({"packet": -1, **r})
    assert got == _json(gold["stream_events"])
    off = S.OfflineASRSession(_asr(staged, "offline"), vad, punc)
    assert off.send(pcm.astype(np.float32) / 32768) == _json(gold["offline_responses"])
