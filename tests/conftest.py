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
