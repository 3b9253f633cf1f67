#!/usr/bin/env python
"""Per-stage error of the CUDA path against the REFERENCE (its ONNX graphs through its vendored onnxruntime, oracle/_ref), on rows
of the benchmark batch: mel, subsampler output, every encoder block output, CTC logits -- tf32 (benchmarked) and exact-fp32 mode.
Runs on the GPU box; writes a markdown table (committed under profiles/).

  python scripts/stage_errors.py [out.md]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import ort_ref  # noqa: E402
from tensorflowasr_b200 import engine as E, weights as W  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "stage_errors.md")
    md = ort_ref.model_dir("offline")
    ge, re_ = W.import_encoder(os.path.join(md, "encoder.onnx"))
    gc, rc = W.import_ctc_model(os.path.join(md, "ctc_model.onnx"))
    x = bench.synth_batch(1234)[:4]                      # row 0 = tiled speech, rows 1..3 = noise (the benchmark batch's first rows)
    nb = ge.num_blocks
    tap_names = ["melspectrogram/Reshape_2:0", "conv_subsampling/dense/BiasAdd:0"] + \
                [f"conformer_block_{i}/layer_normalization_{5 * i + 4}/add:0" for i in range(nb - 1)]   # (the last block's LN is the graph output)
    th = min(16, os.cpu_count() or 1)
    enc_m = ort_ref.OrtModel(os.path.join(md, "encoder.onnx"), th, taps=tap_names)
    ctc_m = ort_ref.OrtModel(os.path.join(md, "ctc_model.onnx"), th)
    feed = {"inputs": x[..., None]}
    ref = {}
    ref["mel"] = enc_m.run(feed, tap_names[0]).reshape(4, -1, 80)
    ref["sub"] = enc_m.run(feed, tap_names[1])
    for i in range(nb - 1):
        ref[f"enc.{i}"] = enc_m.run(feed, tap_names[2 + i])
    enc_out = enc_m.run(feed)
    ref[f"enc.{nb - 1}"] = enc_out
    ref["logits"] = ctc_m.run({"inputs": enc_out})
    rows = {}
    for prec, name in ((0, "tf32"), (1, "fp32")):
        e = E.Engine(ge, re_, gc, rc, precision=prec, use_cuda_graph=False)
        got = {"mel": e.mel(x).cpu().numpy()}
        taps = e.encode_taps(x).cpu().numpy()
        got["sub"] = taps[0]
        for i in range(nb):
            got[f"enc.{i}"] = taps[1 + i]
        got["logits"] = e.ctc_logits(e.encode(x)).cpu().numpy()
        torch.cuda.synchronize()
        rows[name] = {k: (float(np.abs(got[k] - ref[k]).max()), float(np.sqrt(np.mean((got[k] - ref[k]) ** 2)))) for k in ref}
        rows[name]["argmax"] = float((got["logits"].argmax(-1) == ref["logits"].argmax(-1)).mean())
        e.close()
    lines = ["# Per-stage error of the CUDA path against the reference (ONNX Runtime 1.10 on the shipped graphs)",
             "",
             "Input: rows 0..3 of the benchmark batch `bench.synth_batch(1234)` (row 0 tiled speech, rows 1..3 Gaussian noise), 4 x 10 s.",
             "Reference taps: `melspectrogram/Reshape_2:0`, `conv_subsampling/dense/BiasAdd:0`, `conformer_block_<i>/layer_normalization_<5i+4>/add:0`,",
             "`ctc_model.onnx` output.  Generated on a B200 by `scripts/stage_errors.py`.",
             "",
             "| stage | max abs(ref) | tf32: max abs err | tf32: rms err | fp32 mode: max abs err | fp32 mode: rms err |",
             "|---|---|---|---|---|---|"]
    for k in ref:
        lines.append(f"| {k} | {np.abs(ref[k]).max():.4g} | {rows['tf32'][k][0]:.3e} | {rows['tf32'][k][1]:.3e} | {rows['fp32'][k][0]:.3e} | "
                     f"{rows['fp32'][k][1]:.3e} |")
    lines.append("")
    lines.append(f"Per-frame argmax agreement with the reference over the 4 x 250 frames: tf32 {rows['tf32']['argmax']:.4f}, "
                 f"fp32 mode {rows['fp32']['argmax']:.4f}.")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
