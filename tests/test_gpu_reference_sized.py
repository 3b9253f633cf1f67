"""GPU parity tests AT THE BENCHMARKED SHAPES against the reference itself (run with -m gpu on a B200).

The reference = its shipped ONNX graphs through its vendored onnxruntime 1.10.0 (oracle/_ref, staged by oracle/build_ref.py; it
travels to the GPU box) and its own externals/ctc_decoders C++ (oracle/_ref/libctcdec_ref.so).  Each test runs the CUDA path on
the FULL batch of a BASELINE.json configuration and the reference on a slice of the very same batch (rows are independent:
equal-length batches), then asserts

  * greedy token ids identical row by row -- noise rows included, in BOTH precision modes (the tf32 mode is the benchmarked one);
    in tf32 mode a frame may flip only if the reference itself decides it by less than twice the stated tolerance
    (conftest.assert_ids_match; the flips are counted and printed);
  * logits within the stated tolerances (conftest.py: fp32 mode max 2e-3; tf32 mode rms 1e-2, max 0.25 -- the reference's CTC
    decoder amplifies a 1e-3 encoder perturbation up to 100x on single elements, see profiles/r02_stage_errors.md);
  * config 5: the beam-16 hypotheses of the device decoder equal the reference C++ decoder's on the same probabilities.
"""
import os

import numpy as np
import pytest

from conftest import assert_ids_match, check_logits

pytestmark = pytest.mark.gpu


def _need_ref():
    from oracle import ort_ref
    if not ort_ref.available():
        pytest.skip("oracle/_ref (vendored onnxruntime + ONNX graphs) not staged")
    return ort_ref


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


@pytest.fixture(scope="module")
def ref_offline():
    ort_ref = _need_ref()
    return ort_ref.ReferenceASR("offline", threads=min(16, os.cpu_count() or 1))


@pytest.fixture(scope="module", params=[1, 0], ids=["fp32", "tf32"])
def eng(request, offline_weights):
    from tensorflowasr_b200 import engine as E
    ge, re_, gc, rc = offline_weights
    e = E.Engine(ge, re_, gc, rc, precision=request.param, use_cuda_graph=True)
    e.precision = request.param
    yield e
    e.close()


def _greedy(logits_row, blank):
    from oracle import ctc_ref
    return ctc_ref.greedy_decode(logits_row, blank)


def test_config2_bench_batch_vs_reference(eng, ref_offline, torch_mod):
    """BASELINE configs[1]: the exact batch bench.py times (bench.synth_batch(1234): 32 x 10 s, every 4th row tiled speech, the
    rest N(0, 0.1^2) noise).  GPU on all 32 rows; reference on rows 0..7 (2 speech + 6 noise rows)."""
    import bench
    x = bench.synth_batch(1234)
    rows = list(range(8))
    xs = torch_mod.from_numpy(x).cuda()
    ids, lens = eng.recognize(xs)
    logits = eng.ctc_logits(eng.encode(xs))
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    got_logits = logits[rows].cpu().numpy()
    ref_logits = ref_offline.logits(ref_offline.encode(x[rows]))
    assert ref_logits.shape == got_logits.shape == (8, 250, 1332)
    check_logits(got_logits, ref_logits, eng.precision, "config 2 (32 x 10 s), rows 0..7:")
    n_tokens = flips = 0
    for i, r in enumerate(rows):
        flips += assert_ids_match(ids[r, :lens[r]].tolist(), got_logits[i], ref_logits[i], eng.precision, what=f"config 2 row {r}")
        n_tokens += len(_greedy(ref_logits[i], 1331))
    assert n_tokens >= 60                      # the two speech rows carry ~3 x 13 tokens each
    for r in (0, 4):                           # the speech rows decode to exactly the reference's ids in both modes
        assert ids[r, :lens[r]].tolist() == _greedy(ref_logits[r], 1331)
    print(f"config 2, precision {eng.precision}: {flips} low-margin frame(s) of {len(rows) * 250} flipped")


def test_config5_beam16_batch_vs_reference(eng, ref_offline, torch_mod):
    """BASELINE configs[4]: 128 x 5 s, prefix beam 16.  Logits / greedy ids of rows 0..5 against the reference graphs; the device
    beam search against the reference's own C++ decoder (ctc_beam_search_decoder.cpp:18-187) on the same probabilities (fp32
    softmax of the logits this engine produced), rows 0, 1, 8, 9."""
    from oracle import ctc_ref, ctcdec_ref
    if not ctcdec_ref.available():
        pytest.skip("oracle/_ref/libctcdec_ref.so not staged")
    import bench
    L, B = 80000, 128
    x = bench.synth_batch(1237, B=B, L=L, speech_every=8)
    xs = torch_mod.from_numpy(x).cuda()
    logits = eng.ctc_logits(eng.encode(xs))
    gids, glens = eng.recognize(xs)
    gids, glens = gids.cpu().numpy(), glens.cpu().numpy()
    rows = list(range(6))
    ref_logits = ref_offline.logits(ref_offline.encode(x[rows]))
    got = logits[rows].cpu().numpy()
    check_logits(got, ref_logits, eng.precision, "config 5 (128 x 5 s), rows 0..5:")
    for i, r in enumerate(rows):
        assert_ids_match(gids[r, :glens[r]].tolist(), got[i], ref_logits[i], eng.precision, what=f"config 5 row {r}")
    assert gids[0, :glens[0]].tolist() == _greedy(ref_logits[0], 1331)
    ids, lens, scores = eng.ctc_beam(logits, 16)
    ids, lens, scores = ids.cpu().numpy(), lens.cpu().numpy(), scores.cpu().numpy()
    for r in (0, 1, 8, 9):
        lg = logits[r].cpu().numpy()
        probs = ctc_ref.softmax(lg.astype(np.float32)).astype(np.float32)
        ref = ctcdec_ref.beam_search(probs.astype(np.float64), 16)
        n = len(ref)
        got_h = [ids[r, k, :lens[r, k]].tolist() for k in range(16) if lens[r, k] >= 0]
        assert len(got_h) == n
        assert got_h[0] == ref[0][1]                                   # the best hypothesis always
        for k in range(n):
            if got_h[k] != ref[k][1]:                                   # neighbours may swap only when their scores tie in float
                assert abs(ref[k][0] - scores[r, k]) < 1e-3 and sorted(map(tuple, got_h)) == sorted(tuple(h[1]) for h in ref)
        np.testing.assert_allclose(scores[r, :n], [h[0] for h in ref], atol=2e-3)


def test_config3_streaming_batch_vs_reference(streaming_weights, torch_mod):
    """BASELINE configs[2]: StreamingConformerCTC, 64 x 30 s = 3840 independent 8000-sample chunks + the global CTC decoder over
    780 frames.  GPU on the full batch (tf32 and fp32); the reference's streaming graphs on rows 0 (speech) and 1 (noise):
    60 chunks each through encoder.onnx, one ctc_model.onnx call over the concatenated 780 frames (test_asr.py:116-165)."""
    ort_ref = _need_ref()
    from tensorflowasr_b200 import engine as E
    import bench
    sd = ort_ref.model_dir("streaming")
    if sd is None:
        pytest.skip("streaming ONNX graphs not staged")
    th = min(16, os.cpu_count() or 1)
    se = ort_ref.OrtModel(os.path.join(sd, "encoder.onnx"), th)
    sc = ort_ref.OrtModel(os.path.join(sd, "ctc_model.onnx"), th)
    L, B = 480000, 64
    x = bench.synth_batch(1235, B=B, L=L, speech_every=16)
    ref_logits = []
    for r in (0, 1):
        enc = se.run({"inputs": x[r].reshape(60, 8000, 1)})             # [60, 13, 256]
        ref_logits.append(sc.run({"inputs": enc.reshape(1, 780, 256)})[0])
    ge, re_, gc, rc = streaming_weights
    xs = torch_mod.from_numpy(x).cuda()
    for prec in (1, 0):
        e = E.Engine(ge, re_, gc, rc, precision=prec, chunk_samples=8000)
        ids, lens = e.recognize(xs)
        ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
        logits = e.ctc_logits(e.encode(xs[:2])).cpu().numpy()
        check_logits(logits, np.stack(ref_logits), prec, "config 3 (64 x 30 s streaming), rows 0, 1:")
        for i, r in enumerate((0, 1)):
            assert_ids_match(ids[r, :lens[r]].tolist(), logits[i], ref_logits[i], prec, what=f"config 3 row {r}")
        assert ids[0, :lens[0]].tolist() == _greedy(ref_logits[0], 1331)
        assert lens[0] >= 60
        e.close()


def test_stage_taps_vs_reference_golden(eng, golden, ref_wav):
    """Subsampler output (conv1 -> conv2 -> linear) against the reference's ONNX tap on the reference wav
    (golden['wav_sub'] = conv_subsampling/dense/BiasAdd:0; values up to ~250), and the last tap against the encoder output."""
    taps = eng.encode_taps(ref_wav[None]).cpu().numpy()
    sub = taps[0, 0]
    assert sub.shape == golden["wav_sub"].shape
    err = float(np.abs(sub - golden["wav_sub"]).max())
    print(f"subsampler output, precision {eng.precision}: max |x - reference| = {err:.3e} (max |x| = {np.abs(golden['wav_sub']).max():.1f})")
    assert err <= (2e-3 if eng.precision == 1 else 8e-2)
    np.testing.assert_allclose(taps[-1, 0], golden["wav_enc"], atol=2e-4 if eng.precision == 1 else 8e-3)


def test_recognize_honours_input_length(eng, ref_wav):
    """b200asr_recognize_lengths == tf.keras.backend.ctc_decode(ctc_output, input_length) of the batched evaluation
    (am_tester.py:34-40): frames at or beyond input_length are not decoded."""
    x = np.stack([np.tile(ref_wav, 2)[:64000], np.tile(ref_wav, 2)[3000:67000]])
    ids, lens = eng.recognize(x)
    logits = eng.ctc_logits(eng.encode(x)).cpu().numpy()
    fl = np.array([40, 100], np.int32)
    ids2, lens2 = eng.recognize(x, frame_lengths=fl)
    ids2, lens2 = ids2.cpu().numpy(), lens2.cpu().numpy()
    for b in range(2):
        want = _greedy(logits[b, :fl[b]], 1331)
        assert ids2[b, :lens2[b]].tolist() == want
        assert (ids2[b, lens2[b]:] == -1).all()
    assert lens2[0] < int(lens[0]) and lens2[1] == int(lens[1])
