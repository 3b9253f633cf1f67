"""Exporter back to the reference's three-file ONNX layout.

The reference exports its Keras models with tf2onnx (test_asr.py:226-242: encoder.onnx / ctc_model.onnx / translator.onnx) and both
deployment surfaces load exactly those files (Inference/PythonInference/asr/src/asr.py:22-25, Inference/CppInference/onnx/src/core/
asr_session.cpp:77-150: input "inputs" (+ "enc" for the translator), output "Identity:0").  This module writes the same three files
from a weight set of THIS package (the `raw` dictionaries of weights.import_encoder / import_ctc_model / import_translator, or
weights.random_model), so a model that lives here -- e.g. one whose weights were edited or re-trained elsewhere -- keeps running under
the reference's onnxruntime deployments.  The graphs are built op by op with a dependency-free protobuf writer (no onnx package):
opset 13, dynamic batch / length axes.

Graph = the reference's forward (conformer_blocks.py), stated with standard ONNX ops:
  encoder      Conv1D STFT (cos / -sin x window kernels, stride hop, SAME_UPPER) -> |X|^2 -> 10 log10(max(p, 1e-10)) - utterance max,
               floor -80 -> mel MatMul -> 2 x Conv2D(3x3, s2, SAME_UPPER) + ReLU -> merge -> Dense -> N x ConformerBlock
  ctc_model    Dense -> N x ConformerBlock -> Dense (logits)
  translator   Gather(embedding) -> N x RBlock (cross attention of LN(x + PE) over "enc") -> Dense
Every weight of `raw` is stored as an initializer under its own key (prefix "b200asr/"), in its own layout; layout changes an ONNX op
needs (OIHW kernels, per-head projections as one matrix) are Transpose / Reshape nodes on the initializer that onnxruntime folds at load
time.  weights.import_* recognise such files and read the initializers back verbatim: export -> import is the identity on `raw`.

tests/test_export.py runs the exported graphs through the reference's own vendored onnxruntime 1.10.0 and compares them with the
shipped graphs on the reference wav.
"""
from __future__ import annotations

import math
import struct
from typing import Dict, List, Optional, Sequence

import numpy as np

PREFIX = "b200asr/"
OPSET = 13


# ------------------------------------------------------------------------------------------------------------ protobuf writer
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _ld(field: int, payload: bytes) -> bytes:
    return _key(field, 2) + _varint(len(payload)) + payload


def _vi(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(v)


def _str(field: int, s: str) -> bytes:
    return _ld(field, s.encode("utf-8"))


_DT = {np.dtype("float32"): 1, np.dtype("int32"): 6, np.dtype("int64"): 7}


def _tensor(name: str, a: np.ndarray) -> bytes:
    a = np.ascontiguousarray(a)
    out = b"".join(_vi(1, int(d)) for d in a.shape) + _vi(2, _DT[a.dtype]) + _str(8, name) + _ld(9, a.tobytes())
    return out


def _attr(name: str, v) -> bytes:
    out = _str(1, name)
    if isinstance(v, float):
        return out + _key(2, 5) + struct.pack("<f", v) + _vi(20, 1)
    if isinstance(v, int):
        return out + _vi(3, v) + _vi(20, 2)
    if isinstance(v, str):
        return out + _ld(4, v.encode()) + _vi(20, 3)
    if isinstance(v, (list, tuple)):
        return out + b"".join(_vi(8, int(i)) for i in v) + _vi(20, 7)
    raise TypeError(f"attribute {name}: {type(v)}")


def _value_info(name: str, elem: int, dims: Sequence) -> bytes:
    shape = b""
    for d in dims:
        shape += _ld(1, _str(2, d) if isinstance(d, str) else _vi(1, int(d)))
    tensor = _vi(1, elem) + _ld(2, shape)
    return _str(1, name) + _ld(2, _ld(1, tensor))


class GraphBuilder:
    """Appends nodes / initializers; `op()` returns the name of the (first) output."""

    def __init__(self, name: str):
        self.name = name
        self.nodes: List[bytes] = []
        self.inits: List[bytes] = []
        self.inputs: List[bytes] = []
        self.outputs: List[bytes] = []
        self._n = 0
        self._consts: Dict[bytes, str] = {}

    def input(self, name: str, elem: int, dims: Sequence):
        self.inputs.append(_value_info(name, elem, dims))
        return name

    def output(self, name: str, elem: int, dims: Sequence):
        self.outputs.append(_value_info(name, elem, dims))

    def weight(self, key: str, a: np.ndarray) -> str:
        name = PREFIX + key
        self.inits.append(_tensor(name, np.asarray(a, dtype=np.float32)))
        return name

    def const(self, a, dtype=None) -> str:
        a = np.asarray(a, dtype=dtype)
        sig = a.dtype.str.encode() + str(a.shape).encode() + a.tobytes()
        if sig not in self._consts:
            self._n += 1
            name = f"const_{self._n}"
            self.inits.append(_tensor(name, a))
            self._consts[sig] = name
        return self._consts[sig]

    def i64(self, *v) -> str:
        return self.const(np.asarray(v, dtype=np.int64))

    def f32(self, v) -> str:
        return self.const(np.asarray(v, dtype=np.float32))

    def op(self, op_type: str, inputs: Sequence[str], out: Optional[str] = None, **attrs) -> str:
        self._n += 1
        out = out or f"{op_type.lower()}_{self._n}"
        node = b"".join(_str(1, i) for i in inputs) + _str(2, out) + _str(3, f"n{self._n}") + _str(4, op_type)
        node += b"".join(_ld(5, _attr(k, v)) for k, v in attrs.items())
        self.nodes.append(node)
        return out

    def serialize(self) -> bytes:
        g = b"".join(_ld(1, n) for n in self.nodes) + _str(2, self.name) + b"".join(_ld(5, t) for t in self.inits)
        g += b"".join(_ld(11, i) for i in self.inputs) + b"".join(_ld(12, o) for o in self.outputs)
        model = _vi(1, 7) + _str(2, "tensorflowasr_b200.onnx_export") + _str(3, "1") + _ld(7, g) + _ld(8, _str(1, "") + _vi(2, OPSET))
        return model


# ------------------------------------------------------------------------------------------------------------ layers
class _Net:
    def __init__(self, g: GraphBuilder, raw: Dict[str, np.ndarray], eps: float):
        self.g, self.raw, self.eps = g, raw, float(eps)
        self._w: Dict[str, str] = {}

    def w(self, key: str) -> str:
        if key not in self._w:
            self._w[key] = self.g.weight(key, self.raw[key])
        return self._w[key]

    def dense(self, x: str, wkey: str, bkey: Optional[str]) -> str:
        y = self.g.op("MatMul", [x, self.w(wkey)])
        return self.g.op("Add", [y, self.w(bkey)]) if bkey else y

    def layer_norm(self, x: str, p: str) -> str:
        g = self.g
        mu = g.op("ReduceMean", [x], axes=[-1], keepdims=1)
        d = g.op("Sub", [x, mu])
        var = g.op("ReduceMean", [g.op("Mul", [d, d])], axes=[-1], keepdims=1)
        y = g.op("Div", [d, g.op("Sqrt", [g.op("Add", [var, g.f32(self.eps)])])])
        return g.op("Add", [g.op("Mul", [y, self.w(p + ".g")]), self.w(p + ".b")])

    def swish(self, x: str) -> str:
        return self.g.op("Mul", [x, self.g.op("Sigmoid", [x])])

    def ff_module(self, x: str, p: str) -> str:
        g = self.g
        h = self.layer_norm(x, p + ".ln")
        h = self.swish(self.dense(h, p + ".w1", p + ".b1"))
        h = self.dense(h, p + ".w2", p + ".b2")
        return g.op("Add", [x, g.op("Mul", [h, g.f32(0.5)])])

    def _heads(self, x: str, wkey: str, H: int, dh: int, transpose_last: bool = False) -> str:
        """x [B, N, D] -> per-head projection [B, H, N, dh] (or [B, H, dh, N])."""
        g = self.g
        D = self.raw[wkey].shape[1]
        w2d = g.op("Reshape", [g.op("Transpose", [self.w(wkey)], perm=[1, 0, 2]), g.i64(D, H * dh)])        # [H, D, dh] -> [D, H*dh]
        y = g.op("Reshape", [g.op("MatMul", [x, w2d]), g.i64(0, 0, H, dh)])
        return g.op("Transpose", [y], perm=[0, 2, 3, 1] if transpose_last else [0, 2, 1, 3])

    def attention(self, x: str, q_in: str, kv_in: str, p: str) -> str:
        """Residual on x; queries from q_in, keys / values from kv_in (multihead_attention.py:151-188: no q/k/v bias, one output bias)."""
        g = self.g
        H, _, dh = self.raw[p + ".wq"].shape
        q = g.op("Div", [self._heads(q_in, p + ".wq", H, dh), g.f32(math.sqrt(dh))])
        kT = self._heads(kv_in, p + ".wk", H, dh, transpose_last=True)
        v = self._heads(kv_in, p + ".wv", H, dh)
        coef = g.op("Softmax", [g.op("MatMul", [q, kT])], axis=-1)
        o = g.op("Reshape", [g.op("Transpose", [g.op("MatMul", [coef, v])], perm=[0, 2, 1, 3]), g.i64(0, 0, H * dh)])
        wo2d = g.op("Reshape", [self.w(p + ".wo"), g.i64(H * dh, -1)])
        return g.op("Add", [x, g.op("Add", [g.op("MatMul", [o, wo2d]), self.w(p + ".bo")])])

    def conv_module(self, x: str, p: str) -> str:
        g = self.g
        D = self.raw[p + ".dw.w"].shape[1]
        K = self.raw[p + ".dw.w"].shape[0]
        y = self.dense(self.layer_norm(x, p + ".ln"), p + ".pw1.w", p + ".pw1.b")
        a = g.op("Slice", [y, g.i64(0), g.i64(D), g.i64(2)])
        b = g.op("Slice", [y, g.i64(D), g.i64(2 * D), g.i64(2)])
        y = g.op("Mul", [a, g.op("Sigmoid", [b])])
        dw = g.op("Unsqueeze", [g.op("Transpose", [self.w(p + ".dw.w")], perm=[1, 0]), g.i64(1)])              # [K, D] -> [D, 1, K]
        y = g.op("Conv", [g.op("Transpose", [y], perm=[0, 2, 1]), dw], group=D, kernel_shape=[K], strides=[1], auto_pad="SAME_UPPER")
        y = g.op("Transpose", [y], perm=[0, 2, 1])
        y = self.dense(y, p + ".pw.w", p + ".pw.b")
        y = g.op("Add", [g.op("Mul", [y, self.w(p + ".bn.scale")]), self.w(p + ".bn.shift")])
        y = self.dense(self.swish(y), p + ".pw2.w", p + ".pw2.b")
        return g.op("Add", [x, y])

    def conformer_block(self, x: str, p: str) -> str:
        x = self.ff_module(x, p + "ffn1")
        xn = self.layer_norm(x, p + "mhsa.ln")
        x = self.attention(x, xn, xn, p + "mhsa")
        x = self.conv_module(x, p + "conv")
        x = self.ff_module(x, p + "ffn2")
        return self.layer_norm(x, p + "ln")


def _dft_kernels(window: np.ndarray, n_dft: int) -> np.ndarray:
    """[2 * (n_dft/2 + 1), 1, n_dft]: cos rows then -sin rows, each times the window (asr/models/layers/backend.py:27-69)."""
    n = np.arange(n_dft, dtype=np.float64)
    wk = np.arange(n_dft // 2 + 1, dtype=np.float64) * 2 * np.pi / float(n_dft)
    win = np.asarray(window, dtype=np.float64)
    re = np.cos(wk[:, None] * n[None, :]) * win[None, :]
    im = -np.sin(wk[:, None] * n[None, :]) * win[None, :]
    return np.concatenate([re, im], 0)[:, None, :].astype(np.float32)


# ------------------------------------------------------------------------------------------------------------ the three files
def export_encoder(geo, raw: Dict[str, np.ndarray], path: str) -> None:
    g = GraphBuilder("encoder")
    net = _Net(g, raw, geo.ln_eps)
    D, nb = geo.dmodel, geo.n_dft // 2 + 1
    x = g.input("inputs", 1, ["batch", "samples", 1])
    x = g.op("Transpose", [x], perm=[0, 2, 1])                                                                   # [B, 1, L]
    net.w("fe.window")                                                                                           # (kept for the importer; the Conv uses the product)
    spec = g.op("Conv", [x, g.const(_dft_kernels(raw["fe.window"], geo.n_dft))], kernel_shape=[geo.n_dft], strides=[geo.hop], auto_pad="SAME_UPPER")
    re = g.op("Slice", [spec, g.i64(0), g.i64(nb), g.i64(1)])
    im = g.op("Slice", [spec, g.i64(nb), g.i64(2 * nb), g.i64(1)])
    p = g.op("Transpose", [g.op("Add", [g.op("Mul", [re, re]), g.op("Mul", [im, im])])], perm=[0, 2, 1])          # [B, T, 513]
    db = g.op("Div", [g.op("Mul", [g.op("Log", [g.op("Max", [p, g.f32(1e-10)])]), g.f32(10.0)]), g.f32(np.log(np.float32(10.0)))])
    db = g.op("Max", [g.op("Sub", [db, g.op("ReduceMax", [db], axes=[1, 2], keepdims=1)]), g.f32(-80.0)])
    mel = g.op("MatMul", [db, net.w("fe.mel")])                                                                  # [B, T, n_mels]
    y = g.op("Unsqueeze", [mel, g.i64(1)])                                                                       # NCHW [B, 1, T, F]
    for name in ("sub.conv1", "sub.conv2"):
        w = g.op("Transpose", [net.w(name + ".w")], perm=[3, 2, 0, 1])                                           # HWIO -> OIHW
        y = g.op("Relu", [g.op("Conv", [y, w, net.w(name + ".b")], kernel_shape=[3, 3], strides=[2, 2], auto_pad="SAME_UPPER")])
    f2 = raw["sub.lin.w"].shape[0] // D
    y = g.op("Reshape", [g.op("Transpose", [y], perm=[0, 2, 3, 1]), g.i64(0, 0, f2 * D)])                         # merge_two_last_dims
    y = net.dense(y, "sub.lin.w", "sub.lin.b")
    for i in range(geo.num_blocks):
        y = net.conformer_block(y, f"enc.{i}.")
    g.op("Identity", [y], out="Identity:0")
    g.output("Identity:0", 1, ["batch", "frames", D])
    with open(path, "wb") as f:
        f.write(g.serialize())


def export_ctc_model(geo, raw: Dict[str, np.ndarray], path: str) -> None:
    g = GraphBuilder("ctc_model")
    net = _Net(g, raw, geo.ln_eps)
    x = g.input("inputs", 1, ["batch", "frames", geo.dmodel])
    y = net.dense(x, "ctc.proj.w", "ctc.proj.b")
    for i in range(geo.num_blocks):
        y = net.conformer_block(y, f"ctc.blk{i}.")
    y = net.dense(y, "ctc.fc.w", "ctc.fc.b")
    g.op("Identity", [y], out="Identity:0")
    g.output("Identity:0", 1, ["batch", "frames", geo.vocab])
    with open(path, "wb") as f:
        f.write(g.serialize())


def export_translator(geo, raw: Dict[str, np.ndarray], path: str, max_tokens: int = 512) -> None:
    from . import weights as W
    g = GraphBuilder("translator")
    net = _Net(g, raw, geo.ln_eps)
    D = geo.dmodel
    ids = g.input("inputs", 6, ["batch", "tokens"])
    enc = g.input("enc", 1, ["batch", "frames", D])
    x = g.op("Gather", [net.w("tr.emb"), ids], axis=0)                                                            # [B, U, D]
    n_tok = g.op("Gather", [g.op("Shape", [ids]), g.i64(1)], axis=0)                                              # [1] int64
    pe = g.op("Slice", [g.const(W.translator_positional_encoding(max_tokens, D).astype(np.float32)), g.i64(0), n_tok, g.i64(0)])
    for i in range(geo.num_blocks):
        p = f"tr.{i}."
        x = net.ff_module(x, p + "ffn1")
        q_in = net.layer_norm(g.op("Add", [x, pe]), p + "mhsa.ln")
        x = net.attention(x, q_in, enc, p + "mhsa")
        x = net.conv_module(x, p + "conv")
        x = net.ff_module(x, p + "ffn2")
        x = net.layer_norm(x, p + "ln")
    y = net.dense(x, "tr.fc.w", "tr.fc.b")
    g.op("Identity", [y], out="Identity:0")
    g.output("Identity:0", 1, ["batch", "tokens", geo.vocab])
    with open(path, "wb") as f:
        f.write(g.serialize())


def export_model_dir(out_dir: str, enc=None, ctc=None, translator=None) -> None:
    """Write encoder.onnx / ctc_model.onnx / translator.onnx (each argument a (geometry, raw) pair or None) into `out_dir`: the
    directory layout Inference/PythonInference/asr/models/<kind>/ and Inference/CppInference expect."""
    import os
    os.makedirs(out_dir, exist_ok=True)
    if enc is not None:
        export_encoder(enc[0], enc[1], os.path.join(out_dir, "encoder.onnx"))
    if ctc is not None:
        export_ctc_model(ctc[0], ctc[1], os.path.join(out_dir, "ctc_model.onnx"))
    if translator is not None:
        export_translator(translator[0], translator[1], os.path.join(out_dir, "translator.onnx"))


def main(argv=None) -> int:
    """python -m tensorflowasr_b200.onnx_export IN_DIR OUT_DIR: read encoder.onnx / ctc_model.onnx (/ translator.onnx) of a deployment
    directory (the reference's tf2onnx files or files written here) and write them again through this exporter."""
    import argparse
    import os
    from . import weights as W
    ap = argparse.ArgumentParser(prog="python -m tensorflowasr_b200.onnx_export", description=main.__doc__)
    ap.add_argument("in_dir")
    ap.add_argument("out_dir")
    a = ap.parse_args(argv)
    enc = W.import_encoder(os.path.join(a.in_dir, "encoder.onnx"))
    ctc = W.import_ctc_model(os.path.join(a.in_dir, "ctc_model.onnx"))
    tr_path = os.path.join(a.in_dir, "translator.onnx")
    tr = W.import_translator(tr_path) if os.path.isfile(tr_path) else None
    export_model_dir(a.out_dir, enc, ctc, tr)
    print("wrote", ", ".join(sorted(f for f in os.listdir(a.out_dir) if f.endswith(".onnx"))), "to", a.out_dir)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
