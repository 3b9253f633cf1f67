"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, weight blob layout,
ONNX reader, shape arithmetic.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from conftest import ROOT
from tensorflowasr_b200 import engine as E, weights as W, onnx_reader


def test_library_exports_every_declared_symbol():
    lib = E.load_library()
    header = open(os.path.join(ROOT, "include", "b200asr.h")).read()
    declared = set(re.findall(r"B200ASR_API\s+[\w\s\*]+?\b(b200asr_\w+)\s*\(", header))
    assert {"b200asr_create", "b200asr_encode", "b200asr_ctc_logits", "b200asr_ctc_greedy", "b200asr_ctc_beam",
            "b200asr_recognize", "b200asr_recognize_host", "b200asr_destroy"} <= declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200asr.h but not exported"
    assert lib.b200asr_abi_version() == 1


def test_config_struct_matches_header():
    header = open(os.path.join(ROOT, "include", "b200asr.h")).read()
    end = header.index("} b200asr_config;")
    body = header[header.rindex("typedef struct {", 0, end):end]
    names = []
    for line in body.splitlines():
        line = line.split("/*")[0].strip()
        m = re.match(r"(int32_t|float)\s+(.*);", line)
        if m:
            for n in m.group(2).split(","):
                names.append(n.strip().split("[")[0])
    assert names == [f[0] for f in E.Config._fields_]
    assert ctypes.sizeof(E.Config) == 4 * (len(names) - 1) + 4 * 4 == 4 * 25      # 25 32-bit words: the ABI-1 size, translator fields carved out of `reserved`


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ge, re_, gc, rc = W.random_model(0, num_blocks=1)
    with pytest.raises(RuntimeError):
        E.Engine(ge, re_, gc, rc)
    # and the raw C entry point reports an error instead of silently falling back
    lib = E.load_library()
    blob = W.pack_blob(W.device_tensors(ge, re_, gc, rc))
    cfg = E.Config()
    cfg.abi_version = 1
    h = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(blob, len(blob))
    rc_ = lib.b200asr_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc_ != 0 and not h.value
    assert b"CUDA" in lib.b200asr_last_error(None)


def test_blob_layout_roundtrip():
    ge, re_, gc, rc = W.random_model(1, num_blocks=1)
    t = W.device_tensors(ge, re_, gc, rc)
    blob = W.pack_blob(t)
    assert blob[:8] == b"B2ASRW01"
    n = struct.unpack("<I", blob[8:12])[0]
    assert n == len(t)
    for i, (name, arr) in enumerate(t.items()):
        ent = blob[16 + 64 * i:16 + 64 * (i + 1)]
        assert ent[:48].rstrip(b"\0").decode() == name
        off, numel = struct.unpack("<QQ", ent[48:])
        assert off % 128 == 0 and numel == arr.size
        np.testing.assert_array_equal(np.frombuffer(blob, np.float32, numel, off), arr.ravel())


def test_device_packing_matches_reference_math():
    """K-major packing, GLU interleave, BN fold and the 1/sqrt(dh) fold reproduce the raw-layout computation."""
    from oracle import conformer_ref as cr
    ge, re_, gc, rc = W.random_model(2, num_blocks=1)
    t = W.device_tensors(ge, re_, gc, rc)
    D = ge.dmodel
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, D))
    p = "enc.0.conv"
    y_ref = x @ re_[p + ".pw1.w"] + re_[p + ".pw1.b"]
    glu_ref = y_ref[:, :D] * cr.sigmoid(y_ref[:, D:])
    y = x @ t[p + ".pw1.w"].T + t[p + ".pw1.b"]
    np.testing.assert_allclose(y[:, 0::2] * cr.sigmoid(y[:, 1::2]), glu_ref, atol=1e-5)
    z_ref = (x @ re_[p + ".pw.w"] + re_[p + ".pw.b"]) * re_[p + ".bn.scale"] + re_[p + ".bn.shift"]
    np.testing.assert_allclose(x @ t[p + ".pw.w"].T + t[p + ".pw.b"], z_ref, atol=1e-5)
    m = "enc.0.mhsa"
    q_ref = np.einsum("ni,hio->nho", x, re_[m + ".wq"]) / np.sqrt(ge.head_size)
    qkv = x @ t[m + ".wqkv"].T
    np.testing.assert_allclose(qkv[:, :D].reshape(3, ge.num_heads, ge.head_size), q_ref, atol=1e-5)
    o = rng.standard_normal((3, ge.num_heads, ge.head_size))
    np.testing.assert_allclose(o.reshape(3, -1) @ t[m + ".wo"].T, np.einsum("nhi,hio->no", o, re_[m + ".wo"]), atol=1e-5)
    w2 = t["sub.conv2.w"].reshape(D, 3, 3, D)
    np.testing.assert_array_equal(w2[5, 1, 2, :], re_["sub.conv2.w"][1, 2, :, 5])


def test_onnx_reader_on_reference_model(offline_weights):
    from oracle import ort_ref
    g = onnx_reader.load_graph(os.path.join(ort_ref.model_dir("offline"), "ctc_model.onnx"))
    assert g.inputs == ["inputs"] and g.outputs == ["Identity:0"]
    assert g.initializers["fully_connected/Tensordot/ReadVariableOp:0"].shape == (144, 1332)
    assert sum(n.op_type == "Softmax" for n in g.nodes) == 1


def test_same_padding_arithmetic():
    from oracle.conformer_ref import tf_same_pad
    assert tf_same_pad(160000, 1024, 160) == (1000, 432, 432)
    assert tf_same_pad(67263, 1024, 160) == (421, 480, 481)
    assert tf_same_pad(1000, 3, 2) == (500, 0, 1) and tf_same_pad(421, 3, 2) == (211, 1, 1)
    assert tf_same_pad(250, 32, 1) == (250, 15, 16)


def test_warp_fft_phases_on_host(tmp_path):
    """csrc/stft_warp.cuh (the 32 x 32 four-step FFT of the STFT kernel) compiled for the HOST and run lane by lane against a
    direct double-precision DFT: tests/host/stft_warp_host.cu exits 0 when both frames' power spectra agree to 2e-5 of the
    frame maximum (padding on either side and the odd-frame-count case included)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "stft_warp_host")
    src = os.path.join(ROOT, "tests", "host", "stft_warp_host.cu")
    r = subprocess.run([nvcc, "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "max_rel_err" in r.stdout


def test_chunk_weight_packing_layout():
    """ChunkConformer blob: K-major GEMM operands, q/k/v biases concatenated with 1/sqrt(dh) folded into the q part, heads padded to a
    multiple of 4 classes with a never-winning bias, every tensor-core operand an exact tf32 number."""
    from tensorflowasr_b200 import chunk_model as CM
    _, fe_raw, _, _ = W.random_model(0, num_blocks=1)
    geo = CM.ChunkGeometry(enc_blocks=2, helper_blocks=1, txt_classes=33)
    raw = CM.random_chunk_weights(3, fe_raw, geo)
    dev = CM.chunk_device_tensors(raw, geo)
    D, H, dh = geo.dmodel, geo.num_heads, geo.head_size
    assert dev["enc.1.mhsa.wqkv"].shape == (3 * H * dh, D) and dev["enc.1.mhsa.bqkv"].shape == (3 * H * dh,)
    np.testing.assert_allclose(dev["enc.0.mhsa.bqkv"][:H * dh], raw["enc.0.mhsa.bq"].reshape(-1) / np.sqrt(np.float32(dh)), rtol=1e-6)
    np.testing.assert_array_equal(dev["enc.0.mhsa.bqkv"][H * dh:2 * H * dh], raw["enc.0.mhsa.bk"].reshape(-1))
    # row h*dh + o of the q block = column (h, o) of the Keras kernel [D, H, dh], scaled
    want = raw["helper.0.mhsa.wq"][:, 2, 5] / np.sqrt(np.float32(dh))
    np.testing.assert_allclose(dev["helper.0.mhsa.wqkv"][2 * dh + 5], want, rtol=2e-3, atol=1e-6)
    assert dev["dec.fc.w"].shape == (36, D) and dev["dec.fc.b"].shape == (36,) and (dev["dec.fc.b"][33:] < -1e29).all()
    assert (dev["dec.fc.w"][33:] == 0).all() and dev["picker.fc.w"].shape == (280, D)
    for k in ("enc.0.ffn1.w1", "sub.conv2.w", "dec.fc.w", "picker.proj.w", "enc.1.mhsa.wqkv"):
        assert (dev[k].view(np.uint32) & 0x1FFF).max() == 0, k                   # exact tf32 numbers
    assert (dev["enc.0.conv.dw.w"].view(np.uint32) & 0x1FFF).max() != 0         # CUDA-core operands stay fp32
    blob = CM.pack_chunk_blob(raw, geo)
    assert blob[:8] == b"B2ASRW01"
    assert ctypes.sizeof(CM.ChunkConfig) == 4 * 22 + 4 * 8


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under tensorflowasr_b200/ (the product) may import or execute oracle/, and bench.py's
    own arm only does so in its cpu_baseline / reference legs (a product path through the oracle would void every parity claim)."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "tensorflowasr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                tree = ast.parse(open(path, encoding="utf-8").read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path
            if f.endswith((".cu", ".cuh", ".h", ".py")):
                text = open(path, encoding="utf-8").read()
                assert "oracle/_ref" not in text and "oracle/" not in text.replace("# oracle", ""), path
    # bench.py: oracle imports only inside the CPU legs
    tree = ast.parse(open(os.path.join(root, "bench.py"), encoding="utf-8").read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, (ast.Import, ast.ImportFrom)) and any("oracle" in (a.name if isinstance(n, ast.Import) else (n.module or "")) for a in n.names)
                   for n in ast.walk(fn))
        if uses:
            assert fn.name in ("cpu_reference_step", "reference_arm", "cpu_baseline_sample", "_ort_threads_best"), fn.name


def test_every_entry_point_of_the_header_is_bound_or_documented():
    """include/b200asr.h cites, for each entry point family, the reference interface it replaces (file:line)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "b200asr.h"), encoding="utf-8").read()
    for cite in ("asr.py", "asr_session.cpp", "ctc_beam_search_decoder", "chunk_conformer_blocks.py", "vad/src/vad.py", "punc_recover.py",
                 "am_tester.py"):
        assert cite in text, cite


def test_fp16_subsampler_operands_are_packed_exactly():
    """tf32 mode: conv2 / subsampling-linear weights travel a second time as IEEE fp16, two per float32 word of the blob; a tf32-rounded
    weight in fp16's normal range converts without any further rounding (same 11-bit significand)."""
    from tensorflowasr_b200 import weights as W
    ge, re_, gc, rc = W.random_model(2, num_blocks=1)
    dev = W.device_tensors(ge, re_, gc, rc, round_tf32=True)
    D = ge.dmodel
    for key, shape in (("sub.conv2.w", (D, 9 * D)), ("sub.lin.w", (D, re_["sub.lin.w"].shape[0]))):
        w32 = dev[key]
        w16 = dev[key + "16"].view(np.float16).reshape(shape)
        assert dev[key + "16"].dtype == np.float32 and dev[key + "16"].size * 2 == w32.size
        normal = np.abs(w32) >= 6.2e-5
        np.testing.assert_array_equal(w16.astype(np.float32)[normal], w32[normal])          # exact where fp16 is normal
        assert np.abs(w16.astype(np.float32) - w32)[~normal].max(initial=0.0) <= 3.0e-8      # subnormal tail: absolute 2^-25
    assert "sub.conv2.w16" not in W.device_tensors(ge, re_, gc, rc, round_tf32=False)        # the exact-fp32 mode keeps fp32 only
