#!/usr/bin/env python
"""CPU study (no GPU needed): how much of the tf32 path's deviation from the fp64 oracle comes from which stage, and what the
operand rounding mode does to it.  Test infrastructure: uses oracle/ (never imported by the product).

The tcgen05 kind::tf32 datapath reads the top 19 bits of each fp32 operand (truncation).  An operand that was rounded to nearest
tf32 beforehand (weights on the host when the blob is packed, activations in the epilogue that produces them) is exact for the
tensor core, so the choice is ours.  This script emulates a GEMM as fp64 accumulation over operands quantised with
  trunc : both operands truncated (what happens to raw fp32 data)
  rn    : both operands rounded to nearest (ties away), the mode the kernels implement when they round
and prints max |x - oracle| per stage on a slice of the reference wav.

  python scripts/tf32_error_study.py [seconds]
"""
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import conformer_ref as cr, ort_ref  # noqa: E402
from tensorflowasr_b200 import weights as W  # noqa: E402


def q_trunc(x):
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return (b & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


def q_rn(x):
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


def q_none(x):
    return np.asarray(x, dtype=np.float64)


class Emu:
    def __init__(self, qa, qw):
        self.qa, self.qw = qa, qw

    def mm(self, a, w):
        return self.qa(a) @ self.qw(w)

    def ff(self, x, raw, p):
        h = cr.layer_norm(x, raw[p + ".ln.g"], raw[p + ".ln.b"])
        h = cr.swish(self.mm(h, raw[p + ".w1"]) + raw[p + ".b1"])
        return x + 0.5 * (self.mm(h, raw[p + ".w2"]) + raw[p + ".b2"])

    def mhsa(self, x, raw, p):
        xn = cr.layer_norm(x, raw[p + ".ln.g"], raw[p + ".ln.b"])
        wq, wk, wv, wo = (raw[p + s].astype(np.float64) for s in (".wq", ".wk", ".wv", ".wo"))
        H, D, dh = wq.shape
        scale = np.float32(1.0 / np.sqrt(np.float32(dh)))
        q = np.stack([self.mm(xn, wq[h] * scale) for h in range(H)], 2)     # [B, T, H, dh]
        k = np.stack([self.mm(xn, wk[h]) for h in range(H)], 2)
        v = np.stack([self.mm(xn, wv[h]) for h in range(H)], 2)
        s = np.einsum("bnho,bmho->bhnm", self.qa(q), self.qa(k))
        s = s - s.max(-1, keepdims=True)
        e = np.exp(s)
        e = q_trunc(e)                                # P is written to TMEM truncated; the row sum uses the truncated values
        o = np.einsum("bhnm,bmhi->bnhi", e, self.qa(v)) / e.sum(-1).transpose(0, 2, 1)[..., None]
        B, T = x.shape[:2]
        out = self.mm(o.reshape(B, T, H * dh), wo.reshape(H * dh, D)) + raw[p + ".bo"]
        return x + out

    def conv(self, x, raw, p):
        y = cr.layer_norm(x, raw[p + ".ln.g"], raw[p + ".ln.b"])
        y = self.mm(y, raw[p + ".pw1.w"]) + raw[p + ".pw1.b"]
        D = x.shape[-1]
        y = y[..., :D] * cr.sigmoid(y[..., D:])
        K = raw[p + ".dw.w"].shape[0]
        _, pl, pr = cr.tf_same_pad(x.shape[1], K, 1)
        y = cr.depthwise_conv1d(y, raw[p + ".dw.w"], pl, pr)
        y = self.mm(y, raw[p + ".pw.w"] * raw[p + ".bn.scale"][None, :]) + raw[p + ".pw.b"] * raw[p + ".bn.scale"] + raw[p + ".bn.shift"]
        y = cr.swish(y)
        return x + self.mm(y, raw[p + ".pw2.w"]) + raw[p + ".pw2.b"]

    def block(self, x, raw, p):
        x = self.ff(x, raw, p + "ffn1")
        x = self.mhsa(x, raw, p + "mhsa")
        x = self.conv(x, raw, p + "conv")
        x = self.ff(x, raw, p + "ffn2")
        return cr.layer_norm(x, raw[p + "ln.g"], raw[p + "ln.b"])

    def conv2(self, x, w, b):
        B, Hh, Ww, Cin = x.shape
        Ho, pt, pb = cr.tf_same_pad(Hh, 3, 2)
        Wo, pl, pr = cr.tf_same_pad(Ww, 3, 2)
        xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
        out = np.zeros((B, Ho, Wo, w.shape[-1]))
        for i in range(3):
            for j in range(3):
                patch = xp[:, i:i + (Ho - 1) * 2 + 1:2, j:j + (Wo - 1) * 2 + 1:2, :]
                out += self.mm(patch, w[i, j])
        return out + b

    def forward(self, wav, re_, nb, rc, stages):
        mel = cr.melspectrogram(wav, re_)
        x = np.maximum(cr.conv2d_same(mel[..., None], re_["sub.conv1.w"], re_["sub.conv1.b"]), 0)      # conv1 is fp32 CUDA cores
        x = np.maximum(self.conv2(x, re_["sub.conv2.w"], re_["sub.conv2.b"]), 0)
        B, T2, F2, D = x.shape
        stages["conv2"] = x
        x = self.mm(x.reshape(B, T2, F2 * D), re_["sub.lin.w"]) + re_["sub.lin.b"]
        stages["sub"] = x
        for i in range(nb):
            x = self.block(x, re_, f"enc.{i}.")
            stages[f"enc.{i}"] = x
        y = self.mm(x, rc["ctc.proj.w"]) + rc["ctc.proj.b"]
        y = self.block(y, rc, "ctc.blk0.")
        stages["ctc.blk0"] = y
        stages["logits"] = self.mm(y, rc["ctc.fc.w"]) + rc["ctc.fc.b"]
        return stages


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    md = ort_ref.model_dir("offline")
    ge, re_ = W.import_encoder(os.path.join(md, "encoder.onnx"))
    gc, rc = W.import_ctc_model(os.path.join(md, "ctc_model.onnx"))
    w = wave.open(os.path.join(ROOT, "tests", "golden", "BAC009S0764W0121.wav"))
    x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32)[None, :int(secs * 16000)] / 32768
    ref = Emu(q_none, q_none).forward(x, re_, ge.num_blocks, rc, {})
    rows = {}
    for name, (qa, qw) in {"trunc/trunc": (q_trunc, q_trunc), "rn(act)/trunc(w)": (q_rn, q_trunc), "trunc(act)/rn(w)": (q_trunc, q_rn),
                           "rn/rn": (q_rn, q_rn)}.items():
        st = Emu(qa, qw).forward(x, re_, ge.num_blocks, rc, {})
        rows[name] = {k: float(np.abs(st[k] - ref[k]).max()) for k in ref}
        amax = (st["logits"].argmax(-1) == ref["logits"].argmax(-1)).mean()
        rows[name]["argmax agree"] = float(amax)
    keys = list(ref.keys()) + ["argmax agree"]
    print("| stage | max |ref| | " + " | ".join(rows) + " |")
    print("|---|---|" + "---|" * len(rows))
    for k in keys:
        mag = f"{np.abs(ref[k]).max():.3g}" if k in ref else ""
        print(f"| {k} | {mag} | " + " | ".join(f"{rows[n][k]:.3e}" for n in rows) + " |")


if __name__ == "__main__":
    main()
