"""Voice-activity model of the reference's session layer: weights from `vad.onnx`, device engine, reference-shaped wrapper.

Reference: Inference/PythonInference/vad/src/vad.py:11-28 (an onnxruntime session over vad/models/vad.onnx, input [1, N, 80] float32 =
N frames of 80 samples of the 8 kHz signal, output [1, N, 1] logits; callers threshold at 0).  The graph (90 K parameters):
Dense 80 -> Dense 80 + ReLU -> causal Conv1D k=5 + ReLU -> Dense 80 + ReLU -> LayerNorm(eps 1e-3) -> causal Conv1D k=5 + ReLU ->
Dense 80 + ReLU -> Dense 1.

Here: `import_vad` reads the initialisers by walking the graph (no names hard-coded), `VADEngine` is the ctypes binding of
`b200asr_vad_create / b200asr_vad_infer` (csrc/vad_engine.cu: every layer is a launch of the exact-fp32 CUDA-core GEMM -- the causal
convolutions as GEMMs over overlapping rows of a front-padded frame buffer), and `VAD` keeps the reference class's `inference(wav)`.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List

import numpy as np

from . import onnx_reader as R
from . import weights as W

VAD_FRAME = 80          # samples per frame at 8 kHz (10 ms)
VAD_TAPS = 5            # causal Conv1D kernel size
VAD_TENSORS = ("d0.w", "d0.b", "d1.w", "d1.b", "c0.w", "c0.b", "d2.w", "d2.b", "ln.g", "ln.b", "c1.w", "c1.b", "d3.w", "d3.b", "d4.w", "d4.b")


def import_vad(path: str) -> Dict[str, np.ndarray]:
    """vad.onnx -> {'d0.w': [80, 80] (in, out), 'd0.b', ..., 'c0.w': [out, in, taps], ..., 'ln.g', 'ln.b', 'ln.eps', 'd4.w': [80, 1]}."""
    g = R.load_graph(path)
    cons = g.consumers()
    init = g.initializers

    def follow(name: str, ops: List[str]):
        """The first consumer chain name -> ops[0] -> ops[1] ...; returns the last node."""
        node = None
        for op in ops:
            nxt = [n for n in cons.get(name, []) if n.op_type == op]
            if not nxt:
                raise ValueError(f"vad.onnx: expected {op} after {name}")
            node = nxt[0]
            name = node.outputs[0]
        return node

    def const_input(node) -> np.ndarray:
        c = [init[i] for i in node.inputs if i in init]
        if len(c) != 1:
            raise ValueError(f"vad.onnx: {node.op_type} without exactly one constant input")
        return np.asarray(c[0], dtype=np.float32)

    raw: Dict[str, np.ndarray] = {}
    nd = nc = 0
    for n in g.nodes:
        if n.op_type == "MatMul":
            raw[f"d{nd}.w"] = np.asarray(init[n.inputs[1]], dtype=np.float32)                 # [in, out]
            raw[f"d{nd}.b"] = const_input(follow(n.outputs[0], ["Reshape", "Add"])).reshape(-1)
            nd += 1
        elif n.op_type == "Conv":
            w = np.asarray(init[n.inputs[1]], dtype=np.float32)                                # [out, in, 1, taps]
            raw[f"c{nc}.w"] = w[:, :, 0, :]
            raw[f"c{nc}.b"] = const_input(follow(n.outputs[0], ["Squeeze", "Add"])).reshape(-1)
            nc += 1
        elif n.op_type == "BatchNormalization":                                                # tf2onnx's LayerNormalization lowering
            mul = follow(n.outputs[0], ["Reshape", "Mul"])
            raw["ln.g"] = const_input(mul).reshape(-1)
            raw["ln.b"] = const_input(follow(mul.outputs[0], ["Add"])).reshape(-1)
            raw["ln.eps"] = np.float32(n.attrs.get("epsilon", 1e-3))
    if nd != 5 or nc != 2 or "ln.g" not in raw:
        raise ValueError(f"vad.onnx: unexpected graph ({nd} dense, {nc} conv layers)")
    if raw["c0.w"].shape[2] != VAD_TAPS or raw["d0.w"].shape != (VAD_FRAME, VAD_FRAME) or raw["d4.w"].shape != (VAD_FRAME, 1):
        raise ValueError("vad.onnx: unexpected layer shapes")
    return raw


def vad_device_tensors(raw: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Device layout: every GEMM operand [N, K] K-major; a causal conv is a GEMM with K = taps * 80 (k = tap * 80 + in); the single
    output unit of the last layer padded to 4 columns (the CUDA-core GEMM writes float4s)."""
    out: Dict[str, np.ndarray] = {}
    for i in range(4):
        out[f"d{i}.w"] = np.ascontiguousarray(raw[f"d{i}.w"].T)
        out[f"d{i}.b"] = raw[f"d{i}.b"]
    for i in range(2):
        w = raw[f"c{i}.w"]                                                                     # [out, in, taps]
        out[f"c{i}.w"] = np.ascontiguousarray(w.transpose(0, 2, 1).reshape(w.shape[0], -1))    # [out, taps * in]
        out[f"c{i}.b"] = raw[f"c{i}.b"]
    out["ln.g"], out["ln.b"] = raw["ln.g"], raw["ln.b"]
    w4 = np.zeros((4, VAD_FRAME), np.float32)
    w4[0] = raw["d4.w"][:, 0]
    b4 = np.zeros((4,), np.float32)
    b4[0] = raw["d4.b"][0]
    out["d4.w"], out["d4.b"] = w4, b4
    return {k: np.ascontiguousarray(out[k], dtype=np.float32) for k in VAD_TENSORS}


class VADEngine:
    """One b200asr VAD handle on one GPU (b200asr_vad_create)."""

    def __init__(self, raw: Dict[str, np.ndarray], device: int = 0):
        import torch
        from . import engine as E
        if not torch.cuda.is_available():
            raise RuntimeError("VADEngine needs a CUDA device (no CPU fallback)")
        self.lib = E.load_library()
        self.device = int(device)
        blob = W.pack_blob(vad_device_tensors(raw))
        self._blob = blob
        h = ctypes.c_void_p()
        rc = self.lib.b200asr_vad_create(blob, len(blob), ctypes.c_float(float(raw.get("ln.eps", 1e-3))), self.device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("b200asr_vad_create: " + self.lib.b200asr_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200asr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def infer(self, wav, stride: int = 1):
        """wav: [B, N * 80 * stride] float32 CUDA tensor; stride 2 takes every second sample (16 kHz in, the reference's `wav[::2]`).
        Returns logits [B, N] (float32, same device)."""
        import torch
        if wav.dim() != 2 or wav.dtype != torch.float32 or not wav.is_cuda or not wav.is_contiguous():
            raise ValueError("VADEngine.infer: expected a contiguous [B, samples] float32 CUDA tensor")
        B, L = wav.shape
        if L % (VAD_FRAME * stride) != 0 or L == 0:
            raise ValueError(f"VADEngine.infer: {L} samples is not a whole number of {VAD_FRAME * stride}-sample frames")
        N = L // (VAD_FRAME * stride)
        out = torch.empty((B, N), device=wav.device, dtype=torch.float32)
        rc = self.lib.b200asr_vad_infer(self._h, wav.data_ptr(), B, N, int(stride), out.data_ptr(),
                                        torch.cuda.current_stream(wav.device).cuda_stream)
        if rc != 0:
            raise RuntimeError("b200asr_vad_infer: " + self.lib.b200asr_last_error(self._h).decode())
        return out


class VAD:
    """Mirror of Inference/PythonInference/vad/src/vad.py: `VAD(config).inference(wav)` with wav [1, N, 80] -> logits [1, N, 1]."""

    def __init__(self, config=None, model_path: str = "./vad/models/vad.onnx", device: int = 0):
        self.config = config
        self.model_path = model_path
        self.device = device
        self.compile()

    def compile(self):
        self.model = VADEngine(import_vad(self.model_path), self.device)

    def inference(self, wav):
        import torch
        x = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32).reshape(wav.shape[0], -1)).to(f"cuda:{self.device}")
        return self.model.infer(x).cpu().numpy()[..., None]
