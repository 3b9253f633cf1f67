"""CPU tests that PIN the oracle: the NumPy restatement (oracle/conformer_ref.py, oracle/ctc_ref.py) against
(a) committed golden vectors generated from the reference itself (tests/golden/make_golden.py) and, when the
reference binaries are staged, (b) the reference run live (vendored onnxruntime / externals/ctc_decoders C++)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_IDS
from oracle import conformer_ref as cr, ctc_ref, ctcdec_ref, ort_ref


def test_importer_geometry(offline_weights):
    ge, re_, gc, rc = offline_weights
    assert (ge.dmodel, ge.num_blocks, ge.num_heads, ge.head_size, ge.kernel_size, ge.ff_dim) == (144, 13, 4, 36, 32, 576)
    assert (gc.num_blocks, gc.vocab) == (1, 1332)
    assert re_["fe.mel"].shape == (513, 80)
    np.testing.assert_allclose(re_["fe.mel"].sum(0), 1.0, atol=1e-5)          # L1-normalised filters
    assert re_["enc.0.mhsa.wq"].shape == (4, 144, 36) and re_["enc.0.mhsa.wo"].shape == (4, 36, 144)
    assert re_["sub.conv2.w"].shape == (3, 3, 144, 144)


def test_importer_streaming_geometry(streaming_weights):
    ge, re_, gc, rc = streaming_weights
    assert (ge.dmodel, ge.num_blocks, ge.num_heads, ge.head_size, ge.kernel_size) == (256, 4, 4, 64, 5)
    assert gc.kernel_size == 32 and gc.vocab == 1332


def test_window_and_mel_restatement(offline_weights):
    """The synthetic-weight helpers reproduce what the reference bakes into its graphs."""
    from tensorflowasr_b200 import weights as W
    _, re_, _, _ = offline_weights
    np.testing.assert_allclose(W.hann_periodic(1024), re_["fe.window"], atol=1e-6)
    np.testing.assert_allclose(W.slaney_mel_l1(), re_["fe.mel"], atol=2e-6)


def test_encoder_stages_vs_golden(offline_weights, golden, ref_wav):
    ge, re_, gc, rc = offline_weights
    taps = {}
    enc = cr.encoder_forward(ref_wav[None], re_, ge.num_blocks, taps=taps)
    np.testing.assert_allclose(taps["mel"][0], golden["wav_mel"], atol=2e-3)        # dB values in [-80, 0]
    np.testing.assert_allclose(taps["sub"][0], golden["wav_sub"], atol=2e-3, rtol=1e-5)
    np.testing.assert_allclose(enc[0], golden["wav_enc"], atol=2e-4)
    logits = cr.ctc_forward(enc, rc, gc.num_blocks)[0]
    np.testing.assert_allclose(logits[golden["wav_logit_frames"]], golden["wav_logits"], atol=2e-3)
    assert (logits.argmax(-1) == golden["wav_argmax"]).all()
    assert ctc_ref.greedy_decode(logits, 1331) == GOLDEN_IDS == golden["wav_ids"].tolist()


def test_encoder_noise_batch_vs_golden(offline_weights, golden, noise_2x2s):
    ge, re_, gc, rc = offline_weights
    enc = cr.encoder_forward(noise_2x2s, re_, ge.num_blocks)
    np.testing.assert_allclose(enc, golden["noise_enc"], atol=3e-4)
    logits = cr.ctc_forward(enc, rc, 1)
    assert (logits.argmax(-1) == golden["noise_argmax"]).all()


def test_dense_dft_equals_fft_form(offline_weights):
    """power_spectrogram (FFT) is the same sum as the reference's dense DFT-kernel convolution."""
    _, re_, _, _ = offline_weights
    x = (np.random.default_rng(3).standard_normal((1, 3000)) * 0.1)
    a = cr.power_spectrogram(x, re_["fe.window"])
    b = cr.power_spectrogram_dense(x, re_["fe.window"].astype(np.float64))
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)


def test_streaming_vs_golden(streaming_weights, golden, ref_wav):
    """Block streaming: chunks are encoded independently (test_asr.py:116-135), one CTC decode over all frames."""
    ge, re_, gc, rc = streaming_weights
    parts = [cr.encoder_forward(ref_wav[None, s:s + 8000], re_, ge.num_blocks) for s in range(0, len(ref_wav), 8000)]
    enc = np.concatenate(parts, axis=1)
    np.testing.assert_allclose(enc[0], golden["stream_enc"], atol=5e-4)
    logits = cr.ctc_forward(enc, rc, 1)[0]
    assert (logits.argmax(-1) == golden["stream_argmax"]).all()
    assert ctc_ref.greedy_decode(logits, 1331) == golden["stream_ids"].tolist() == GOLDEN_IDS


@pytest.mark.skipif(not ort_ref.available(), reason="reference onnxruntime not staged")
def test_oracle_vs_live_reference_random_input(offline_weights):
    ge, re_, gc, rc = offline_weights
    ref = ort_ref.ReferenceASR("offline", 1)
    x = (np.random.default_rng(11).standard_normal((1, 12345)) * 0.05).astype(np.float32)
    enc_ref = ref.encode(x)
    enc = cr.encoder_forward(x, re_, ge.num_blocks)
    np.testing.assert_allclose(enc, enc_ref, atol=3e-4)
    np.testing.assert_allclose(cr.ctc_forward(enc_ref, rc, 1), ref.logits(enc_ref), atol=2e-3)


# ------------------------------------------------------------------------------------------------- CTC decoders
def test_greedy_edge_cases():
    V = 5
    assert ctc_ref.greedy_decode(np.zeros((0, V)), V - 1) == []
    y = np.full((4, V), -5.0)
    y[:, V - 1] = 3.0                       # all blank
    assert ctc_ref.greedy_decode(y, V - 1) == []
    y = np.eye(V)[[1, 1, 4, 1, 2, 2, 4, 4, 0]]
    assert ctc_ref.greedy_decode(y, V - 1) == [1, 1, 2, 0]
    y = np.zeros((3, V))                    # ties -> first index wins
    assert ctc_ref.greedy_decode(y, V - 1) == [0]


def test_beam_golden(golden, offline_weights, ref_wav):
    ge, re_, gc, rc = offline_weights
    logits = cr.ctc_forward(cr.encoder_forward(ref_wav[None], re_, ge.num_blocks), rc, 1)[0]
    probs = ctc_ref.softmax(logits)
    res = ctc_ref.beam_search(probs, 4)
    assert [r[1] for r in res] == golden["wav_beam_ids"].tolist()
    np.testing.assert_allclose([r[0] for r in res], golden["wav_beam_scores"], atol=2e-4)
    assert res[0][1] == GOLDEN_IDS
    assert ctc_ref.beam_search(probs, 1)[0][1] == GOLDEN_IDS       # beam 1 == greedy invariant on this utterance


@pytest.mark.skipif(not ctcdec_ref.available(), reason="reference ctc_decoders not built")
@pytest.mark.parametrize("beam,cutoff_prob,cutoff_top_n", [(1, 1.0, 40), (4, 1.0, 40), (8, 0.99, 40), (8, 0.999, 10)])
def test_beam_restatement_vs_reference_cpp(beam, cutoff_prob, cutoff_top_n):
    rng = np.random.default_rng(beam * 100 + cutoff_top_n)
    for T, V, scale in ((1, 7, 1.0), (25, 30, 3.0), (40, 60, 6.0)):
        probs = ctc_ref.softmax(rng.standard_normal((T, V)) * scale)
        mine = ctc_ref.beam_search(probs, beam, cutoff_prob=cutoff_prob, cutoff_top_n=cutoff_top_n)
        ref = ctcdec_ref.beam_search(probs, beam, cutoff_prob, cutoff_top_n)
        assert [m[1] for m in mine] == [r[1] for r in ref]
        np.testing.assert_allclose([m[0] for m in mine], [r[0] for r in ref], atol=1e-4)
        assert ctcdec_ref.greedy(probs) == ctc_ref.greedy_decode(probs, V - 1)


def test_translator_oracle_pinned_to_reference_golden(golden):
    """oracle.conformer_ref.translator_forward (embedding -> RBlocks with cross attention over the encoder states -> Dense) against the
    reference's translator.onnx run by its own onnxruntime on the reference wav (tests/golden/make_golden.py): per-position argmax
    (= the characters of '甚至出现交易几乎停制的情况' + </S>) identical, logits within 1e-3."""
    import os
    from oracle import ort_ref, conformer_ref as cr
    from tensorflowasr_b200 import weights as W
    md = ort_ref.model_dir("offline")
    if md is None or not os.path.isfile(os.path.join(md, "translator.onnx")):
        pytest.skip("translator.onnx not staged")
    gt, rt = W.import_translator(os.path.join(md, "translator.onnx"))
    assert (gt.num_blocks, gt.dmodel, gt.vocab) == (2, 144, 9160)
    out = cr.translator_forward(golden["wav_tr_in"][None], golden["wav_enc"][None], rt, gt.num_blocks)
    assert (out[0].argmax(-1) == golden["wav_tr_argmax"]).all()
    np.testing.assert_allclose(out[0, :4], golden["wav_tr_logits_rows"], atol=1e-3)
    pe = cr.positional_encoding(7, 6)
    np.testing.assert_allclose(pe[3], [np.sin(3.0), np.cos(3.0), np.sin(3 / 10000 ** (2 / 6)), np.cos(3 / 10000 ** (2 / 6)),
                                       np.sin(3 / 10000 ** (4 / 6)), np.cos(3 / 10000 ** (4 / 6))], rtol=1e-5)
    np.testing.assert_allclose(W.translator_positional_encoding(7, 6), pe, atol=1e-7)
