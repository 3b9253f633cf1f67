"""Punctuation model of the reference's session layer: weights from `punc.onnx`, device engine, reference-shaped wrapper.

Reference: Inference/PythonInference/punc_recover/src/punc_recover.py:12-62 (`Punc.punc_recover(txt)`: an onnxruntime session over
punc_recover/models/punc.onnx with inputs token ids [1, U] (<S> + characters + </S>), a padding mask and a sinusoidal table [1, 1024, 64];
output class probabilities [1, U, 32]; a punctuation mark is inserted after a character when argmax > 1 and max >= 0.65).  The graph
("PuncTransformer", 2.5 MB):
    x = embedding[ids] * 8 + PE;  x = ELU(Dense(x))
    3 x { y = EncoderLayer(x);  x = ReLU(causal Conv1D_k3(y)) + x }
    h = Dense_64(Dense_768(x));  h = EncoderLayer(h);  h = EncoderLayer(h);  softmax(Dense_32(h))
    EncoderLayer(x): x1 = LN(x + MHA(x));  LN(x1 + Dense(ReLU(Dense(x1))))      (8 heads of 8, q/k/v/out biases, LN eps 1e-6)

Here: `import_punc` reads the initialisers (layer stems from the names, roles from the numeric order tf.keras gives them),
`PuncEngine` is the ctypes binding of `b200asr_punc_create / b200asr_punc_infer` (csrc/punc_engine.cu: exact-fp32 CUDA-core GEMMs, the
fp32 attention kernel, LayerNorm; one sentence per call as in the reference, so the padding mask is empty), `Punc` keeps the
reference class's `punc_recover(txt)`.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
import re
from typing import Dict, List

import numpy as np

from . import onnx_reader as R
from . import weights as W

PUNC_LAYERS = 5            # 3 inside the convolutional encoder + 2 after the bottleneck
PUNC_CONVS = 3
PUNC_MAX_TOKENS = 1024     # model_config.pe_input


def _num(tag: str) -> int:
    m = re.search(r"_(\d+)$", tag)
    return int(m.group(1)) if m else 0


def import_punc(path: str) -> Dict[str, np.ndarray]:
    """punc.onnx -> {'emb' [V, 64], 'in.w' [64, 64] (in, out), 'in.b', 'l{i}.{q,k,v,o,f1,f2}.{w,b}', 'l{i}.ln{1,2}.{g,b}', 'c{i}.w' [out, in, 3],
    'c{i}.b', 'up.w' [64, 768], 'up.b', 'down.w', 'down.b', 'out.w' [64, 32], 'out.b', 'ln.eps', 'emb.scale'}."""
    g = R.load_graph(path)
    init = g.initializers
    cons = g.consumers()
    raw: Dict[str, np.ndarray] = {}

    def dense(stem: str):
        w = [k for k in init if re.fullmatch(re.escape(stem) + r"/(Tensordot|MatMul)/ReadVariableOp:0", k)]
        b = [k for k in init if re.fullmatch(re.escape(stem) + r"/BiasAdd/ReadVariableOp:0", k)]
        if len(w) != 1 or len(b) != 1:
            raise ValueError(f"punc.onnx: dense layer {stem} not found")
        return np.asarray(init[w[0]], np.float32), np.asarray(init[b[0]], np.float32)

    emb = [k for k in init if re.search(r"embedding/embedding_lookup/\d+:0$", k)]
    if len(emb) != 1:
        raise ValueError("punc.onnx: embedding table not found")
    raw["emb"] = np.asarray(init[emb[0]], np.float32)
    raw["in.w"], raw["in.b"] = dense("encoder/dense")
    # encoder layers: stems "<scope>encoder_layer[_i]" in index order
    stems = sorted({m.group(1) for k in init for m in [re.match(r"^((?:encoder/)?encoder_layer(?:_\d+)?)/", k)] if m}, key=lambda s: _num(s.split("/")[-1]))
    if len(stems) != PUNC_LAYERS:
        raise ValueError(f"punc.onnx: {len(stems)} encoder layers (expected {PUNC_LAYERS})")
    for i, st in enumerate(stems):
        keys = [k for k in init if k.startswith(st + "/")]
        att = sorted({m.group(1) for k in keys for m in [re.match(r"^(.*/multi_head_attention(?:_\d+)?/dense(?:_\d+)?)/", k)] if m}, key=lambda s: _num(s.split("/")[-1]))
        ffn = sorted({m.group(1) for k in keys for m in [re.match(r"^(.*/sequential(?:_\d+)?/dense(?:_\d+)?)/", k)] if m}, key=lambda s: _num(s.split("/")[-1]))
        lns = sorted({m.group(1) for k in keys for m in [re.match(r"^(.*/layer_normalization(?:_\d+)?)/batchnorm/", k)] if m}, key=lambda s: _num(s.split("/")[-1]))
        if len(att) != 4 or len(ffn) != 2 or len(lns) != 2:
            raise ValueError(f"punc.onnx: unexpected structure of {st}")
        for tag, stem in zip(("q", "k", "v", "o"), att):          # tf.keras creates wq, wk, wv, dense in this order
            raw[f"l{i}.{tag}.w"], raw[f"l{i}.{tag}.b"] = dense(stem)
        for tag, stem in zip(("f1", "f2"), ffn):
            raw[f"l{i}.{tag}.w"], raw[f"l{i}.{tag}.b"] = dense(stem)
        for tag, stem in zip(("ln1", "ln2"), lns):
            raw[f"l{i}.{tag}.g"] = np.asarray(init[stem + "/batchnorm/mul/ReadVariableOp:0"], np.float32)
            raw[f"l{i}.{tag}.b"] = np.asarray(init[stem + "/batchnorm/ReadVariableOp:0"], np.float32)
    eps = [v for k, v in init.items() if k.endswith("layer_normalization/batchnorm/add/y:0")]
    raw["ln.eps"] = np.float32(eps[0]) if eps else np.float32(1e-6)
    # causal convolutions: weight from the Conv node, bias from the Add behind its Squeeze
    convs = [n for n in g.nodes if n.op_type == "Conv"]
    if len(convs) != PUNC_CONVS:
        raise ValueError(f"punc.onnx: {len(convs)} Conv nodes (expected {PUNC_CONVS})")
    for i, n in enumerate(convs):
        w = np.asarray(init[n.inputs[1]], np.float32)                                   # [out, in, 1, taps]
        raw[f"c{i}.w"] = w[:, :, 0, :]
        nxt = n
        for op in ("Squeeze", "Add"):
            nxt = [c for c in cons.get(nxt.outputs[0], []) if c.op_type == op][0]
        raw[f"c{i}.b"] = np.asarray([init[x] for x in nxt.inputs if x in init][0], np.float32).reshape(-1)
    # bottleneck + head: the three dense layers outside any encoder layer, by output width
    outside = sorted({m.group(1) for k in init for m in [re.match(r"^((?:time_distributed/)?dense_\d+)/", k)] if m}, key=lambda s: _num(s.split("/")[-1]))
    if len(outside) != 3:
        raise ValueError("punc.onnx: expected three dense layers behind the encoder")
    (raw["up.w"], raw["up.b"]), (raw["down.w"], raw["down.b"]), (raw["out.w"], raw["out.b"]) = (dense(s) for s in outside)
    sq = [v for k, v in init.items() if k.endswith("encoder/Sqrt:0")]
    raw["emb.scale"] = np.float32(sq[0]) if sq else np.float32(np.sqrt(raw["emb"].shape[1]))
    return raw


def punc_positional_encoding(rows: int = PUNC_MAX_TOKENS, d_model: int = 64) -> np.ndarray:
    """Punc.get_pos_encoding (punc_recover.py:20-35): float64 angles, sin on even / cos on odd columns, cast to float32."""
    pos = np.arange(rows)[:, np.newaxis]
    i = np.arange(d_model)[np.newaxis, :]
    ang = pos * (1 / np.power(10000, (2 * (i // 2)) / np.float32(d_model)))
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return np.array(ang, "float32")


def punc_device_tensors(raw: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Device layout: GEMM operands [N, K] K-major; q | k | v stacked into one [192, 64] operand with 1/sqrt(head size) folded into the q
    rows (and bias); causal conv as a GEMM with K = 3 * 64 (k = tap * 64 + in); embedding pre-scaled; the positional table."""
    D = raw["emb"].shape[1]
    H = 8
    scale = np.float32(1.0 / np.sqrt(D // H))
    out: Dict[str, np.ndarray] = {"emb": raw["emb"] * raw["emb.scale"], "pe": punc_positional_encoding(PUNC_MAX_TOKENS, D),
                                  "in.w": raw["in.w"].T, "in.b": raw["in.b"]}
    for i in range(PUNC_LAYERS):
        p = f"l{i}."
        out[p + "qkv.w"] = np.concatenate([raw[p + "q.w"].T * scale, raw[p + "k.w"].T, raw[p + "v.w"].T], 0)
        out[p + "qkv.b"] = np.concatenate([raw[p + "q.b"] * scale, raw[p + "k.b"], raw[p + "v.b"]], 0)
        for t in ("o", "f1", "f2"):
            out[p + t + ".w"], out[p + t + ".b"] = raw[p + t + ".w"].T, raw[p + t + ".b"]
        for t in ("ln1", "ln2"):
            out[p + t + ".g"], out[p + t + ".b"] = raw[p + t + ".g"], raw[p + t + ".b"]
    for i in range(PUNC_CONVS):
        w = raw[f"c{i}.w"]
        out[f"c{i}.w"], out[f"c{i}.b"] = w.transpose(0, 2, 1).reshape(w.shape[0], -1), raw[f"c{i}.b"]
    for t in ("up", "down", "out"):
        out[t + ".w"], out[t + ".b"] = raw[t + ".w"].T, raw[t + ".b"]
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


class PuncEngine:
    """One b200asr punctuation handle on one GPU (b200asr_punc_create)."""

    def __init__(self, raw: Dict[str, np.ndarray], device: int = 0):
        import torch
        from . import engine as E
        if not torch.cuda.is_available():
            raise RuntimeError("PuncEngine needs a CUDA device (no CPU fallback)")
        self.lib = E.load_library()
        self.device = int(device)
        self.vocab = int(raw["emb"].shape[0])
        self.classes = int(raw["out.b"].shape[0])
        blob = W.pack_blob(punc_device_tensors(raw))
        h = ctypes.c_void_p()
        rc = self.lib.b200asr_punc_create(blob, len(blob), ctypes.c_float(float(raw["ln.eps"])), self.device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("b200asr_punc_create: " + self.lib.b200asr_last_error(None).decode())
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

    def infer(self, ids):
        """ids: [U] int32 CUDA tensor (one sentence: <S> characters </S>, no padding) -> class probabilities [U, classes] (float32)."""
        import torch
        if ids.dim() != 1 or ids.dtype != torch.int32 or not ids.is_cuda or not ids.is_contiguous():
            raise ValueError("PuncEngine.infer: expected a contiguous [U] int32 CUDA tensor")
        U = ids.shape[0]
        out = torch.empty((U, self.classes), device=ids.device, dtype=torch.float32)
        rc = self.lib.b200asr_punc_infer(self._h, ids.data_ptr(), U, out.data_ptr(), torch.cuda.current_stream(ids.device).cuda_stream)
        if rc != 0:
            raise RuntimeError("b200asr_punc_infer: " + self.lib.b200asr_last_error(self._h).decode())
        return out


class Punc:
    """Mirror of punc_recover/src/punc_recover.py:12-62: `Punc(config).punc_recover(txt)` -> list of characters with marks inserted."""

    def __init__(self, config, model_path: str = "./punc_recover/models/punc.onnx", device: int = 0):
        from .asr import TextFeaturizer
        self.running_config = config["running_config"]
        self.model_config = config["model_config"]
        self.vocab_featurizer = TextFeaturizer(config["punc_vocab"])
        self.bd_featurizer = TextFeaturizer(config["punc_biaodian"])
        self.model_path = model_path
        self.device = device
        self.compile()

    def compile(self):
        self.model = PuncEngine(import_punc(self.model_path), self.device)

    def probabilities(self, txt) -> np.ndarray:
        import torch
        x = [self.vocab_featurizer.startid()] + self.vocab_featurizer.extract(txt) + [self.vocab_featurizer.endid()]
        if 0 in x:
            raise ValueError("token id 0 is the padding id: a single sentence never contains it")
        ids = torch.tensor(x, dtype=torch.int32, device=f"cuda:{self.device}")
        return self.model.infer(ids).cpu().numpy()

    def punc_recover(self, txt) -> List[str]:
        pred = self.probabilities(txt)[1:-1]
        new_txt = []
        for t, b in zip(txt, pred):
            new_txt.append(t)
            if b.argmax() > 1 and b.max() >= 0.65:
                new_txt.append(self.bd_featurizer.vocab_array[b.argmax()])
        return new_txt
