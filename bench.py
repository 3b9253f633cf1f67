#!/usr/bin/env python
"""Benchmark of the B200 Conformer-CTC hot path (BASELINE.json metric: audio frames/sec, 16 kHz, 10 s utterances).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision tf32|fp32]

A step = one pass of the hot path (wav -> mel -> ConformerCTC(S) encoder -> CTC decoder -> greedy ids) over one
batch of 32 x 10 s synthetic 16 kHz utterances per GPU (BASELINE.json configs[1]).  1 frame = one 10 ms hop.
`value` = frames/s with the waveforms resident in HBM (device-timed, CUDA events, max over ranks);
`e2e`   = the same through the C-ABI call that takes HOST buffers (pinned H2D of the batch + D2H of ids inside the
          timed region).
N > 1 (torchrun): utterances shard across ranks (weak scaling, one replica per GPU); the only collective is the
NCCL all_gather of the decoded ids + lengths, inside the timed region.
`--impl reference` times the reference's own CPU deployment path (shipped ONNX graphs through its vendored
onnxruntime 1.10.0, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH, SECONDS, SR, HOP = 32, 10, 16000, 160
L = SECONDS * SR
FRAMES_PER_UTT = L // HOP
METRIC = "audio frames/sec (16 kHz, 10 s utts)"


def synth_batch(seed: int, B: int = BATCH, L: int = L, speech_every: int = 4) -> np.ndarray:
    """wav ~ N(0, 0.1^2) clipped to [-1, 1]; every `speech_every`-th row is tiled speech so the decoder emits tokens."""
    rng = np.random.default_rng(seed)
    x = np.clip(rng.standard_normal((B, L)).astype(np.float32) * 0.1, -1.0, 1.0)
    wav_path = os.path.join(ROOT, "tests", "golden", "BAC009S0764W0121.wav")
    if os.path.isfile(wav_path):
        import wave
        w = wave.open(wav_path)
        s = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768
        x[::speech_every] = np.tile(s, L // len(s) + 1)[:L]
    return x


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


# ----------------------------------------------------------------------------------------------------------- reference arm
def best_cpu_reference(x):
    """ONNX Runtime 1.10 loses throughput when its intra-op pool is oversubscribed (128 threads ran 10x slower than 16 on the
    GPU box's host).  The reference arm therefore gets the thread count that serves it best: one timed pass per candidate,
    keep the fastest.  Returns (ReferenceASR, threads)."""
    from oracle import ort_ref
    cores = os.cpu_count() or 1
    cands = sorted({min(cores, c) for c in (8, 16, 32, 64, cores)})
    best = None
    for th in cands:
        ref = ort_ref.ReferenceASR("offline", threads=th)
        ref.logits(ref.encode(x[:1]))                      # warm-up (graph optimisation, arena)
        t0 = time.perf_counter()
        ref.logits(ref.encode(x))
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, ref, th)
    return best[1], best[2]


def reference_arm(args, rank: int, world: int):
    """The reference's own CPU implementation of the path (kind 'reference': oracle/_ref ONNX Runtime + shipped graphs)."""
    if rank != 0:
        return
    from oracle import ort_ref, ctc_ref
    line = {"impl": "reference", "metric": METRIC, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": 0}
    if not ort_ref.available():
        line["unavailable"] = "oracle/_ref (vendored onnxruntime + ONNX graphs) was not staged in this checkout"
        print(json.dumps(line))
        return
    sample_b = 4                                         # bounded sample: 4 of the 32 utterances per step
    x = synth_batch(1234)[:sample_b]
    ref, cores = best_cpu_reference(x)

    def step():
        enc = ref.encode(x)
        logits = ref.logits(enc)
        return [ctc_ref.greedy_decode(l, logits.shape[-1] - 1) for l in logits]

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample_b * FRAMES_PER_UTT * args.steps / dt
    line.update({"value": val, "ms_per_step": dt / args.steps * 1e3,
                 "config": {"workload": f"ConformerCTC(S) offline greedy, {sample_b} x 10 s sample of the 32 x 10 s batch, "
                                        "ONNX Runtime 1.10.0 CPU", "global_batch": sample_b, "seq_len": L},
                 "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "reference",
                                  "sample": f"{sample_b} x 10 s utterances per step, {args.steps} steps, {cores} of {os.cpu_count()} host threads "
                                            "(fastest of 8/16/32/64/all)"},
                 "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def cpu_baseline_sample():
    """Bounded CPU baseline timed beside the GPU run (rank 0, N=1 only)."""
    from oracle import ort_ref, ctc_ref
    if not ort_ref.available():
        return None
    sample_b = 4
    x = synth_batch(1234)[:sample_b]
    ref, cores = best_cpu_reference(x)

    def step():
        logits = ref.logits(ref.encode(x))
        return [ctc_ref.greedy_decode(l, logits.shape[-1] - 1) for l in logits]

    step()
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 40):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": sample_b * FRAMES_PER_UTT * n / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": f"{n} passes over {sample_b} x 10 s utterances (ONNX Runtime 1.10.0, {cores} of {os.cpu_count()} host threads: "
                      "the fastest of 8/16/32/64/all)"}


# ----------------------------------------------------------------------------------------------------------- own arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from tensorflowasr_b200 import engine as E, weights as W
    from oracle import ort_ref  # only to locate the staged reference weights; nothing from oracle/ is on the timed path

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version (and, with NCCL_DEBUG=INFO, much more) on stdout while the communicator comes up; stdout must
        # carry only the JSON line, so file descriptor 1 points at stderr until the first collective has completed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    md = ort_ref.model_dir("offline")
    if md is not None:
        ge, re_ = W.import_encoder(os.path.join(md, "encoder.onnx"))
        gc, rc = W.import_ctc_model(os.path.join(md, "ctc_model.onnx"))
        weights_desc = "reference-trained ConformerCTC(S) weights (shipped ONNX)"
    else:
        ge, re_, gc, rc = W.random_model(0, num_blocks=13)
        weights_desc = "random-init ConformerCTC(S) architecture"
    prec = E.PRECISION_TF32 if args.precision == "tf32" else E.PRECISION_FP32
    eng = E.Engine(ge, re_, gc, rc, device=local_rank, precision=prec, use_cuda_graph=True)
    eng.reserve(BATCH, L)
    Tp = eng.out_frames(L)

    # inputs: NROT distinct batches (> L2 together) rotated so no step re-reads a cached waveform
    NROT = 8
    host = [torch.from_numpy(synth_batch(1234 + 97 * rank + i)).pin_memory() for i in range(NROT)]
    dev = [h.cuda(non_blocking=True) for h in host]
    ids = torch.empty((BATCH, Tp), device="cuda", dtype=torch.int32)
    lens = torch.empty((BATCH,), device="cuda", dtype=torch.int32)
    gather_ids = [torch.empty_like(ids) for _ in range(world)] if world > 1 else None
    gather_lens = [torch.empty_like(lens) for _ in range(world)] if world > 1 else None
    stream = torch.cuda.Stream()

    def step(i):
        eng.recognize(dev[i % NROT], ids, lens)
        if world > 1:
            dist.all_gather(gather_ids, ids)
            dist.all_gather(gather_lens, lens)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, NROT)):     # at least one pass over every rotated input: each has its own captured CUDA graph
            step(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        launches = eng.launch_count - launches0

        # end to end through the host-buffer C-ABI entry point
        hid = torch.empty((BATCH, Tp), dtype=torch.int32).pin_memory()
        hlen = torch.empty((BATCH,), dtype=torch.int32).pin_memory()
        dev_in = torch.empty_like(dev[0])

        def e2e_step(i):
            if world == 1:
                eng.recognize_host(host[i % NROT], hid, hlen)          # H2D + compute + D2H + sync inside the C-ABI call
            else:
                dev_in.copy_(host[i % NROT], non_blocking=True)
                eng.recognize(dev_in, ids, lens)
                dist.all_gather(gather_ids, ids)
                dist.all_gather(gather_lens, lens)
                hid.copy_(gather_ids[rank], non_blocking=True)
                hlen.copy_(gather_lens[rank], non_blocking=True)
                torch.cuda.current_stream().synchronize()

        for i in range(args.warmup):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        barrier()
        e2e_sync_s = time.perf_counter() - t0
        e2e_s = e2e_sync_s
        # the same through the two-deep pipeline of the C ABI (b200asr_recognize_host_submit / _collect): every step still
        # copies its own 20.5 MB waveform batch from pinned host memory and reads its ids back, but the H2D of step i+1 overlaps
        # the compute of step i.  (N > 1 builds the same overlap from the Python API, see below.)
        e2e_mode = "synchronous b200asr_recognize_host call per step"
        if world > 1:
            # N > 1: the same two-deep overlap built from the Python API (Engine.recognize on device buffers + NCCL all_gather of the
            # ids): a copy stream brings step i+1's waveforms from pinned host memory into the other device buffer while step i
            # computes; every step still moves its own 20.5 MB in and its gathered ids out
            copy_stream = torch.cuda.Stream()
            dev_in2 = [torch.empty_like(dev[0]) for _ in range(2)]
            hid2 = [hid, torch.empty_like(hid).pin_memory()]
            hlen2 = [hlen, torch.empty_like(hlen).pin_memory()]
            h2d_done = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            done = [torch.cuda.Event() for _ in range(2)]
            for ev in consumed:
                ev.record(stream)

            def submit(i, base):
                sl = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[sl])              # the compute that last read this buffer has finished
                    dev_in2[sl].copy_(host[(base + i) % NROT], non_blocking=True)
                    h2d_done[sl].record(copy_stream)
                stream.wait_event(h2d_done[sl])
                eng.recognize(dev_in2[sl], ids, lens)
                consumed[sl].record(stream)
                dist.all_gather(gather_ids, ids)
                dist.all_gather(gather_lens, lens)
                hid2[sl].copy_(gather_ids[rank], non_blocking=True)
                hlen2[sl].copy_(gather_lens[rank], non_blocking=True)
                done[sl].record(stream)

            def pipelined_n(n, base):
                for i in range(n):
                    if i >= 2:
                        done[i & 1].synchronize()
                    submit(i, base)
                for i in range(max(n - 2, 0), n):
                    done[i & 1].synchronize()

            pipelined_n(max(args.warmup, 2), 0)
            barrier()
            t0 = time.perf_counter()
            pipelined_n(args.steps, 3)
            barrier()
            e2e_s = time.perf_counter() - t0
            e2e_mode = "two-deep pipeline (copy stream + Engine.recognize + NCCL all_gather of ids): H2D of step i+1 under the compute of step i"
        if world == 1:
            hid2 = [hid, torch.empty_like(hid).pin_memory()]
            hlen2 = [hlen, torch.empty_like(hlen).pin_memory()]

            def pipelined(n, base):
                for i in range(n):
                    sl = i & 1
                    if i >= 2:
                        eng.recognize_host_collect(sl)
                    eng.recognize_host_submit(sl, host[(base + i) % NROT], hid2[sl], hlen2[sl])
                for i in range(max(n - 2, 0), n):
                    eng.recognize_host_collect(i & 1)

            pipelined(max(args.warmup, 2), 0)
            barrier()
            t0 = time.perf_counter()
            pipelined(args.steps, 3)
            barrier()
            e2e_s = time.perf_counter() - t0
            e2e_mode = "two-deep pipeline (b200asr_recognize_host_submit/_collect): H2D of step i+1 under the compute of step i"
        sampler.stop_flag.set()
        sampler.join(timeout=2)

        # roofline of the dominant kernel, timed alone with CUDA events on this stream (back-to-back launches through the C ABI)
        roof = None
        if rank == 0:
            eng.recognize(dev[0], ids, lens)
            torch.cuda.synchronize()
            hbm_peak, tf_peak, how = measured_peaks()
            traffic = {}
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram bytes per launch from the committed ncu --set full capture
            if os.path.isfile(tpath):
                traffic = json.load(open(tpath))
            prec = "tf32 (half the bf16 tensor rate)" if args.precision == "tf32" else "fp32 on CUDA cores"

            def stage_line(st, iters=20):
                m, f, b = eng.time_stage(st, BATCH, L, iters=iters)
                d = {"ms_per_launch": round(m, 5), "tflops": round(f / (m * 1e-3) / 1e12, 2), "frac_tensor_peak": round(f / (m * 1e-3) / 1e12 / tf_peak, 4),
                     "gbs": round(b / (m * 1e-3) / 1e9, 1), "frac_hbm_peak": round(b / (m * 1e-3) / 1e9 / hbm_peak, 4),
                     "flops_per_launch": f, "algorithmic_bytes_per_launch": b, "traffic": traffic.get(st)}
                return m, f, b, d

            # the chained FFModule kernel (both GEMMs + swish + residual + LayerNorm; 42 of the 121 launches, ~35 % of the step)
            ms_k, flops, bytes_, _ = stage_line("ffn_chain")
            ach = flops / (ms_k * 1e-3) / 1e12
            roof = {"kernel": "FFModule as one chained tcgen05 kernel, cluster-pair variant (M=%d, 144 -> 576 -> 144, swish, 0.5-residual, "
                              "LayerNorm): 42 of the 121 launches of a step" % (BATCH * Tp),
                    "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({how}); the kernel computes in {prec}",
                    "ms_per_launch": ms_k, "flops_per_launch": flops, "algorithmic_bytes_per_launch": bytes_,
                    "traffic": traffic.get("ffn_chain"),
                    "note": "M = 8000 rows is 63 row tiles for 148 SMs and a tf32 tcgen05.mma (M=128, K=8) costs ~132 cycles for any N <= 144 "
                            "(profiles/r01_ubench_mma_tf32_pacing.txt): the kernel is bound by its MMA instruction count, not by bytes"}
            others = {}
            for st in ("conv2", "stft", "conv1", "qkv", "attention", "dwconv", "sub_linear", "ctc_fc"):
                try:
                    others[st] = stage_line(st)[3]
                except RuntimeError as ex:            # e.g. conv1 has no kernel of its own when it is fused into conv2's
                    others[st] = {"skipped": str(ex)[:160]}
            roof["other_stages"] = others

    t = torch.tensor([ms_total, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    frames = BATCH * FRAMES_PER_UTT * world * args.steps
    if rank == 0:
        value = frames / (ms_total * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
                "rtf": (ms_total * 1e-3) / (BATCH * SECONDS * world * args.steps),
                "config": {"workload": "ConformerCTC(S) 10M offline greedy, batch 32 x 10 s synthetic 16 kHz per GPU "
                                       "(BASELINE.json configs[1])", "weights": weights_desc, "global_batch": BATCH * world,
                           "seq_len": L, "parallelism": f"dp{world} (utterance shard, ids all_gather)",
                           "l2_policy": f"{NROT} distinct input batches rotated ({NROT * BATCH * L * 4 / 1e6:.0f} MB > 126 MB L2); "
                                        "per-step intermediates (369 MB conv1 map) exceed L2",
                           "frame": "10 ms hop (160 samples)"},
                "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": BATCH * L * 4,
                        "d2h_bytes_per_step": BATCH * Tp * 4 + BATCH * 4, "ms_per_step": e2e_ms / args.steps,
                        "mode": e2e_mode, "sync_call_ms_per_step": e2e_sync_s * 1e3 / args.steps},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_sample()
            except Exception as ex:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"error": str(ex)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
