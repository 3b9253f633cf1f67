"""CPU tests of oracle/chunk_conformer_ref.py (parity UNPINNED: no weights, no TensorFlow).  What can
be checked without the reference is the reference's own consistency criterion (test_chunk_asr.py:57,123,139): the streaming
path with state caches reproduces the offline path frame for frame."""
import numpy as np
import pytest

from oracle import chunk_conformer_ref as cc
from tensorflowasr_b200 import weights as W

CFG = dict(cc.CFG, enc_blocks=3, phone_classes=12, txt_classes=30)


@pytest.fixture(scope="module")
def model():
    _, fe_raw, _, _ = W.random_model(0, num_blocks=1)
    return cc.random_chunk_model(7, fe_raw, CFG)


@pytest.fixture(scope="module")
def wav():
    rng = np.random.default_rng(3)
    t = np.arange(2560 * 30) / 16000.0
    return (0.3 * np.sin(2 * np.pi * 440 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.05 * rng.standard_normal(t.size))[None]


def test_chunk_mask_matches_reference_formula():
    """_compute_chunk_mask (:158-176) written out with loops."""
    for n, wf, wb in [(10, 3, 0), (50, 36, 0), (50, 36, 8), (5, 36, 8), (20, 6, 2)]:
        m = cc.chunk_mask(n, wf, wb)
        for i in range(n):
            low = max(i - wf, 0)
            high = min(max(i + wb, 0), n)
            low = low - max(low - n + wb, 0)
            high = high + max(wb - high, 0)
            for j in range(n):
                assert m[i, j] == (not (j < low or j > high))


def test_front_end_streaming_equals_offline(model, wav):
    """Streaming mel frames of chunk i are the offline frames 16i .. 16i+15 (the 1023 zeros the 'valid' mel layer prepends only
    reach frames that the [-16:] slice discards), and the subsampled frames are the offline frames 4i .. 4i+3."""
    off_mel = cc.mel_valid(wav, model)
    off = cc.front_call(wav, model)
    wav_c, sub_c = cc.front_init_caches(1, CFG)
    outs = []
    for i, s in enumerate(range(0, wav.shape[1], 2560)):
        chunk = wav[:, s:s + 2560]
        mel = cc.mel_valid(np.concatenate([wav_c, chunk], 1), model)[:, -16:]
        np.testing.assert_allclose(mel, off_mel[:, 16 * i:16 * i + 16], atol=1e-9)
        o, wav_c, sub_c = cc.front_stream_call(chunk, wav_c, sub_c, model, CFG)
        assert o.shape[1] == 4 and wav_c.shape[1] == 2560 and sub_c.shape[1] == 4
        outs.append(o)
    got = np.concatenate(outs, 1)
    np.testing.assert_allclose(got, off[:, :got.shape[1]], atol=1e-8)
    assert off.shape[1] - got.shape[1] <= 1


def test_encoder_and_picker_streaming_equal_offline(model, wav):
    """win_back = 0 everywhere up to the picker: 36-frame attention cache + 32-frame causal-conv cache reproduce the offline
    band mask / causal convolution exactly."""
    x = cc.encoder_call(cc.front_call(wav, model), model, CFG)
    phone_off, hid_off = cc.ctc_decoder_call(x, model, "picker", CFG["picker_blocks"], 0, CFG)
    caches = cc.init_picker_caches(1, CFG)
    ph, hid = [], []
    for s in range(0, wav.shape[1], 2560):
        v, u, h, caches = cc.picker_stream_predict(wav[:, s:s + 2560], caches, model, CFG)
        assert v.shape[1] == 4 and caches[6].shape[1] == 0
        assert all(c.shape[1] <= 36 for c in caches[2]) and all(c.shape[1] <= 32 for c in caches[3])
        ph.append(v)
        hid.append(h)
    ph, hid = np.concatenate(ph, 1), np.concatenate(hid, 1)
    np.testing.assert_allclose(hid, hid_off[:, :hid.shape[1]], atol=1e-7)
    np.testing.assert_allclose(ph, phone_off[:, :ph.shape[1]], atol=1e-7)


def test_full_stream_equals_predict_on_valid_frames(model, wav):
    """test_chunk_asr.py's loop (picker step -> feature_pick -> decoder step with an 8-frame look-ahead) against predict():
    every 'valid' text frame equals the offline frame; the frames still inside the look-ahead are the 'unvalid' tail."""
    off = cc.predict(wav, model, CFG)
    txt, unvalid, phones = cc.stream_utterance(wav, model, CFG)
    n = txt.shape[1]
    assert n >= 60 and 0 <= off.shape[1] - n <= 8 + 4
    np.testing.assert_allclose(txt, off[:, :n], atol=1e-6)
    assert unvalid.shape[1] == 8


def test_feature_pick_compacts_non_blank_frames():
    h = np.arange(2 * 5 * 3, dtype=float).reshape(2, 5, 3)
    c = np.zeros((2, 5, 4))
    c[0, [0, 2, 3], 1] = 1          # non-blank at frames 0, 2, 3
    c[0, [1, 4], 3] = 1             # blank = class 3
    c[1, :, 3] = 1
    c[1, 4, 0] = 2
    f, cp = cc.feature_pick(h, c, blank=3)
    assert f.shape == (2, 3, 3)
    np.testing.assert_array_equal(f[0], h[0, [0, 2, 3]])
    np.testing.assert_array_equal(f[1, 0], h[1, 4])
    assert (f[1, 1:] == 0).all() and (cp[1, 1:] == 0).all()
