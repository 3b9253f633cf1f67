"""ONNX exporter: the three files onnx_export.py writes, run through the REFERENCE's own vendored onnxruntime 1.10.0,
reproduce the shipped graphs they were exported from; export -> import is the identity on the weights."""
import os

import numpy as np
import pytest

from tensorflowasr_b200 import onnx_export as X
from tensorflowasr_b200 import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ort():
    from oracle import ort_ref
    if not ort_ref.available():
        pytest.skip("oracle/_ref not staged (run oracle/build_ref.py)")
    return ort_ref


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests/golden/offline_golden.npz"))


@pytest.fixture(scope="module")
def wav():
    import wave
    w = wave.open(os.path.join(ROOT, "tests/golden/BAC009S0764W0121.wav"))
    return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768


@pytest.fixture(scope="module")
def exported(ort, tmp_path_factory):
    d = ort.model_dir("offline")
    out = str(tmp_path_factory.mktemp("export"))
    enc = W.import_encoder(os.path.join(d, "encoder.onnx"))
    ctc = W.import_ctc_model(os.path.join(d, "ctc_model.onnx"))
    tr = W.import_translator(os.path.join(d, "translator.onnx")) if os.path.isfile(os.path.join(d, "translator.onnx")) else None
    X.export_model_dir(out, enc, ctc, tr)
    return out, enc, ctc, tr


def test_export_import_is_identity(exported):
    out, enc, ctc, tr = exported
    for fn, name, (geo, raw) in ((W.import_encoder, "encoder.onnx", enc), (W.import_ctc_model, "ctc_model.onnx", ctc)) + \
            (((W.import_translator, "translator.onnx", tr),) if tr else ()):
        geo2, raw2 = fn(os.path.join(out, name))
        assert geo2 == geo
        assert set(raw2) == set(raw)
        for k in raw:
            np.testing.assert_array_equal(np.asarray(raw2[k]), np.asarray(raw[k]), err_msg=k)


def test_exported_graphs_reproduce_the_reference(exported, ort, gold, wav):
    """Same wav through the exported files and through the shipped ones, both on the reference's onnxruntime: encoder states, CTC
    logits and argmax, translator logits and characters."""
    out, enc, ctc, tr = exported
    m = ort.OrtModel(os.path.join(out, "encoder.onnx"), 4)
    e = m.run({"inputs": wav.reshape(1, -1, 1)})
    assert e.shape == (1,) + gold["wav_enc"].shape
    assert np.abs(e[0] - gold["wav_enc"]).max() < 1e-4          # measured 1.8e-6 (fp32 graphs with different op orders)
    c = ort.OrtModel(os.path.join(out, "ctc_model.onnx"), 4)
    logits = c.run({"inputs": gold["wav_enc"][None]})[0]
    assert np.abs(logits[gold["wav_logit_frames"]] - gold["wav_logits"]).max() < 5e-4   # measured 1.1e-5
    np.testing.assert_array_equal(logits.argmax(-1), gold["wav_argmax"])
    logits_e2e = c.run({"inputs": e})[0]
    np.testing.assert_array_equal(logits_e2e.argmax(-1), gold["wav_argmax"])
    if tr is not None and "wav_tr_in" in gold:
        t = ort.OrtModel(os.path.join(out, "translator.onnx"), 4)
        y = t.run({"inputs": gold["wav_tr_in"][None].astype(np.int32), "enc": gold["wav_enc"][None]})[0]
        assert np.abs(y[:4] - gold["wav_tr_logits_rows"]).max() < 5e-4   # measured 4.8e-5
        np.testing.assert_array_equal(y.argmax(-1), gold["wav_tr_argmax"])


def test_exported_graph_is_dynamic_in_batch_and_length(exported, ort):
    out, *_ = exported
    m = ort.OrtModel(os.path.join(out, "encoder.onnx"), 4)
    ref = ort.OrtModel(os.path.join(ort.model_dir("offline"), "encoder.onnx"), 4)
    rng = np.random.default_rng(0)
    for B, L in ((1, 1600), (3, 8000), (2, 12345)):
        x = (rng.standard_normal((B, L, 1)) * 0.1).astype(np.float32)
        a, b = m.run({"inputs": x}), ref.run({"inputs": x})
        assert a.shape == b.shape
        assert np.abs(a - b).max() < 2e-4


def test_streaming_model_exports_too(ort, tmp_path, wav):
    """dmodel 256 / 4 x 64 heads / kernel 5 (the StreamingConformerCTC files): same exporter, same check."""
    ds = ort.model_dir("streaming")
    if ds is None:
        pytest.skip("streaming models not staged")
    X.export_model_dir(str(tmp_path), W.import_encoder(os.path.join(ds, "encoder.onnx")), W.import_ctc_model(os.path.join(ds, "ctc_model.onnx")))
    x = wav[:8000].reshape(1, -1, 1)
    a = ort.OrtModel(str(tmp_path / "encoder.onnx"), 4).run({"inputs": x})
    b = ort.OrtModel(os.path.join(ds, "encoder.onnx"), 4).run({"inputs": x})
    assert a.shape == b.shape and np.abs(a - b).max() < 1e-4
    la = ort.OrtModel(str(tmp_path / "ctc_model.onnx"), 4).run({"inputs": b})
    lb = ort.OrtModel(os.path.join(ds, "ctc_model.onnx"), 4).run({"inputs": b})
    assert np.abs(la - lb).max() < 5e-4


def test_random_model_exports_and_runs(ort, tmp_path):
    """A model that never was an ONNX file (weights.random_model) -> exported -> the reference's onnxruntime == the oracle."""
    from oracle import conformer_ref
    ge, re_, gc, rc = W.random_model(3, num_blocks=2)
    X.export_model_dir(str(tmp_path), (ge, re_), (gc, rc))
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((2, 4000)) * 0.1).astype(np.float32)
    e = ort.OrtModel(str(tmp_path / "encoder.onnx"), 4).run({"inputs": x[..., None]})
    want = conformer_ref.encoder_forward(x, re_, ge.num_blocks)
    assert np.abs(e - want).max() < 2e-3
    lg = ort.OrtModel(str(tmp_path / "ctc_model.onnx"), 4).run({"inputs": e})
    want = conformer_ref.ctc_forward(e, rc, gc.num_blocks)
    assert np.abs(lg - want).max() < 2e-3


def test_command_line_round_trip(ort, tmp_path):
    """python -m tensorflowasr_b200.onnx_export IN OUT on the shipped directory, then once more on its own output: same bytes."""
    import subprocess
    import sys
    d = ort.model_dir("offline")
    a, b = tmp_path / "a", tmp_path / "b"
    for src, dst in ((d, a), (str(a), b)):
        r = subprocess.run([sys.executable, "-m", "tensorflowasr_b200.onnx_export", src, str(dst)], cwd=ROOT, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    for f in ("encoder.onnx", "ctc_model.onnx"):
        assert (a / f).read_bytes() == (b / f).read_bytes()
