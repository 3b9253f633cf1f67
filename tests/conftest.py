import os
import sys
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_IDS = [669, 82, 103, 78, 247, 56, 71, 573, 386, 82, 30, 213, 496]


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "offline_golden.npz"))


@pytest.fixture(scope="session")
def ref_wav():
    w = wave.open(os.path.join(GOLDEN, "BAC009S0764W0121.wav"))
    return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768


def _models(kind):
    from oracle import ort_ref
    d = ort_ref.model_dir(kind)
    if d is None:
        pytest.skip(f"reference {kind} ONNX models not staged (oracle/build_ref.py)")
    return d


@pytest.fixture(scope="session")
def offline_weights():
    from tensorflowasr_b200 import weights as W
    d = _models("offline")
    ge, re_ = W.import_encoder(os.path.join(d, "encoder.onnx"))
    gc, rc = W.import_ctc_model(os.path.join(d, "ctc_model.onnx"))
    return ge, re_, gc, rc


@pytest.fixture(scope="session")
def streaming_weights():
    from tensorflowasr_b200 import weights as W
    d = _models("streaming")
    ge, re_ = W.import_encoder(os.path.join(d, "encoder.onnx"))
    gc, rc = W.import_ctc_model(os.path.join(d, "ctc_model.onnx"))
    return ge, re_, gc, rc


@pytest.fixture(scope="session")
def noise_2x2s():
    rng = np.random.default_rng(7)
    return np.clip(rng.standard_normal((2, 32000)).astype(np.float32) * 0.1, -1, 1)


# ---------------------------------------------------------------------------------------------------------- tolerances
# Stated tolerances of the two precision modes against the reference (fp32 ONNX graphs) / the fp64 oracle.
#   fp32 mode (CUDA-core GEMMs): encoder states 2e-4, logits 2e-3 -- and ids / per-frame argmax identical, no exceptions.
#   tf32 mode (tcgen05, every operand rounded to nearest tf32 by its producer): encoder states 8e-3 (measured 3.7e-3 on values
#     up to 8, i.e. 2^-11 relative: the rounding floor of one tf32 operand), logits rms 1e-2 (measured 2.4e-3 .. 3.3e-3) and max 0.25 (measured up to 0.19 on noise rows).  The max is an outlier
#     bound: the reference's CTC decoder amplifies a 1e-3 perturbation of the encoder output up to 100x on single elements (conv
#     module x12, final LayerNorm gain 8; scripts/tf32_error_study.py, profiles/r02_stage_errors.md), so no single-pass tf32
#     arithmetic can hold the 2e-2 that the rms suggests.  Greedy ids are identical to the reference wherever the reference's own
#     top-2 logit margin exceeds 2 x the max tolerance; a frame inside that margin may legitimately flip (assert_ids_match).
TOL_ENC = {1: 2e-4, 0: 8e-3}
TOL_LOGITS_MAX = {1: 2e-3, 0: 0.25}
TOL_LOGITS_RMS = {1: 2e-4, 0: 1e-2}


def check_logits(got, ref, precision, what=""):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    d = got - ref
    mx, rms = float(np.abs(d).max()), float(np.sqrt(np.mean(d * d)))
    print(f"{what} precision {precision}: logits max |err| {mx:.3e}, rms {rms:.3e}")
    assert mx <= TOL_LOGITS_MAX[precision], (what, mx)
    assert rms <= TOL_LOGITS_RMS[precision], (what, rms)
    return mx, rms


def assert_ids_match(got_ids, got_logits_row, ref_logits_row, precision, blank=None, what=""):
    """got_ids == greedy(ref_logits_row): exactly in fp32 mode.  In tf32 mode a frame whose top-2 margin in the reference is
    below 2 x TOL_LOGITS_MAX is undecidable at the stated tolerance: the per-frame argmax of the device logits may differ from the
    reference's on such frames ONLY, and got_ids must be the CTC collapse of the device's own argmax sequence.  Returns the number
    of frames that flipped (0 in the common case)."""
    from oracle import ctc_ref
    ref_logits_row, got_logits_row = np.asarray(ref_logits_row), np.asarray(got_logits_row)
    blank = ref_logits_row.shape[-1] - 1 if blank is None else blank
    want = ctc_ref.greedy_decode(ref_logits_row, blank)
    ga, ra = got_logits_row.argmax(-1), ref_logits_row.argmax(-1)
    flipped = np.nonzero(ga != ra)[0]
    if len(flipped) == 0:
        assert list(got_ids) == want, (what, list(got_ids), want)
        return 0
    assert precision == 0, (what, "exact mode: a per-frame argmax differs from the reference at frames", flipped.tolist())
    top2 = np.sort(ref_logits_row[flipped], axis=-1)[:, -2:]
    margin = top2[:, 1] - top2[:, 0]
    assert (margin < 2 * TOL_LOGITS_MAX[0]).all(), (what, "argmax flipped on a frame the reference decides by more than the tolerance",
                                                    flipped.tolist(), margin.tolist())
    assert list(got_ids) == ctc_ref.greedy_decode(got_logits_row, blank), (what, "ids are not the collapse of the device's own argmax")
    print(f"{what}: {len(flipped)} low-margin frame(s) flipped (reference margins {np.round(margin, 4).tolist()})")
    return len(flipped)
