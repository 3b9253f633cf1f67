"""TEST INFRASTRUCTURE ONLY (oracle) -- ctypes driver for oracle/_ref/libortref.so.

Runs the reference's shipped ONNX graphs through the reference's vendored onnxruntime 1.10.0, i.e. the exact
deployment path of Inference/PythonInference/asr/src/asr.py:22-75 and
Inference/CppInference/onnx/src/core/asr_session.cpp:77-122.  Also able to expose internal tensors ("taps") by
appending graph outputs to a temp copy of the model.
"""
from __future__ import annotations

import ctypes
import os
import tempfile
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_REFERENCE_MODELS = "/root/reference/Inference/PythonInference/asr/models"


def model_dir(kind: str = "offline") -> Optional[str]:
    """Directory holding encoder.onnx / ctc_model.onnx for `kind` in {offline, streaming} (None if absent)."""
    for base in (os.environ.get("B200ASR_MODEL_ROOT", ""), os.path.join(REF_DIR, "models"), _REFERENCE_MODELS):
        if base and os.path.isfile(os.path.join(base, kind, "encoder.onnx")):
            return os.path.join(base, kind)
    return None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "libortref.so")) and model_dir("offline") is not None


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(os.path.join(REF_DIR, "libortref.so"))
        lib.ortref_open.restype = ctypes.c_void_p
        lib.ortref_open.argtypes = [ctypes.c_char_p, ctypes.c_int]
        lib.ortref_close.argtypes = [ctypes.c_void_p]
        lib.ortref_error.restype = ctypes.c_char_p
        lib.ortref_error.argtypes = [ctypes.c_void_p]
        lib.ortref_run.restype = ctypes.c_int
        lib.ortref_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                   ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.POINTER(ctypes.c_int64)),
                                   ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_char_p,
                                   ctypes.POINTER(ctypes.POINTER(ctypes.c_float)), ctypes.POINTER(ctypes.c_int64),
                                   ctypes.POINTER(ctypes.c_int)]
        lib.ortref_free.argtypes = [ctypes.POINTER(ctypes.c_float)]
        _lib = lib
    return _lib


# ---------------------------------------------------------------- protobuf surgery for taps
def _enc_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _dec_varint(buf: bytes, pos: int):
    r = 0
    s = 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << s
        if not b & 0x80:
            return r, pos
        s += 7


def _ld(field: int, payload: bytes) -> bytes:
    return _enc_varint((field << 3) | 2) + _enc_varint(len(payload)) + payload


def _value_info(name: str) -> bytes:
    tensor = _enc_varint((1 << 3) | 0) + _enc_varint(1)          # Tensor.elem_type = FLOAT
    typ = _ld(1, tensor)                                         # TypeProto.tensor_type
    return _ld(1, name.encode()) + _ld(2, typ)                   # ValueInfoProto{name, type}


def add_graph_outputs(model_bytes: bytes, names: Sequence[str]) -> bytes:
    """Return a copy of the serialized ModelProto whose graph lists `names` as extra outputs (GraphProto.output=12)."""
    out = bytearray()
    pos = 0
    n = len(model_bytes)
    while pos < n:
        start = pos
        key, pos = _dec_varint(model_bytes, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            _, pos = _dec_varint(model_bytes, pos)
        elif wt == 1:
            pos += 8
        elif wt == 5:
            pos += 4
        elif wt == 2:
            ln, pos2 = _dec_varint(model_bytes, pos)
            if fno == 7:
                graph = bytes(model_bytes[pos2:pos2 + ln]) + b"".join(_ld(12, _value_info(nm)) for nm in names)
                out += _ld(7, graph)
                pos = pos2 + ln
                continue
            pos = pos2 + ln
        out += model_bytes[start:pos]
    return bytes(out)


class OrtModel:
    """One onnxruntime session over a reference ONNX file."""

    def __init__(self, path: str, threads: int = 1, taps: Sequence[str] = ()):
        lib = _load()
        self._tmp = None
        if taps:
            with open(path, "rb") as f:
                patched = add_graph_outputs(f.read(), taps)
            fd, self._tmp = tempfile.mkstemp(suffix=".onnx")
            with os.fdopen(fd, "wb") as f:
                f.write(patched)
            path = self._tmp
        self._h = lib.ortref_open(path.encode(), int(threads))
        if not self._h:
            raise RuntimeError(f"onnxruntime could not open {path}")

    def close(self):
        if self._h:
            _load().ortref_close(self._h)
            self._h = None
        if self._tmp and os.path.exists(self._tmp):
            os.unlink(self._tmp)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, feeds: Dict[str, np.ndarray], output: str = "Identity:0") -> np.ndarray:
        lib = _load()
        names = list(feeds.keys())
        arrs = [np.ascontiguousarray(feeds[k]) for k in names]
        n = len(arrs)
        c_names = (ctypes.c_char_p * n)(*[k.encode() for k in names])
        c_data = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
        dims_keep = [(ctypes.c_int64 * a.ndim)(*a.shape) for a in arrs]
        c_dims = (ctypes.POINTER(ctypes.c_int64) * n)(*[ctypes.cast(d, ctypes.POINTER(ctypes.c_int64)) for d in dims_keep])
        c_nd = (ctypes.c_int * n)(*[a.ndim for a in arrs])
        dts = []
        for a in arrs:
            if a.dtype == np.float32:
                dts.append(0)
            elif a.dtype == np.int32:
                dts.append(1)
            else:
                raise TypeError(f"unsupported feed dtype {a.dtype}")
        c_dt = (ctypes.c_int * n)(*dts)
        out_ptr = ctypes.POINTER(ctypes.c_float)()
        out_dims = (ctypes.c_int64 * 8)()
        out_nd = ctypes.c_int(0)
        rc = lib.ortref_run(self._h, n, c_names, c_data, c_dims, c_nd, c_dt, output.encode(), ctypes.byref(out_ptr),
                            out_dims, ctypes.byref(out_nd))
        if rc != 0:
            raise RuntimeError("onnxruntime: " + lib.ortref_error(self._h).decode(errors="replace"))
        shape = tuple(int(out_dims[i]) for i in range(out_nd.value))
        count = int(np.prod(shape)) if shape else 1
        res = np.ctypeslib.as_array(out_ptr, shape=(count,)).copy().reshape(shape)
        lib.ortref_free(out_ptr)
        return res


class ReferenceASR:
    """encoder.onnx + ctc_model.onnx exactly as Inference/PythonInference/asr/src/asr.py:34-75 drives them."""

    def __init__(self, kind: str = "offline", threads: int = 1):
        d = model_dir(kind)
        if d is None:
            raise FileNotFoundError("reference ONNX models not found (run oracle/build_ref.py)")
        self.encoder = OrtModel(os.path.join(d, "encoder.onnx"), threads)
        self.ctc = OrtModel(os.path.join(d, "ctc_model.onnx"), threads)

    def encode(self, wav: np.ndarray) -> np.ndarray:
        wav = np.asarray(wav, dtype=np.float32)
        if wav.ndim == 1:
            wav = wav.reshape(1, -1, 1)
        elif wav.ndim == 2:
            wav = wav[..., None]
        return self.encoder.run({"inputs": wav})

    def logits(self, enc: np.ndarray) -> np.ndarray:
        return self.ctc.run({"inputs": np.asarray(enc, dtype=np.float32)})
