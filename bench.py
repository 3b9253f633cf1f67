#!/usr/bin/env python
"""Benchmark of the B200 Conformer-CTC hot path (BASELINE.json metric: audio frames/sec, 16 kHz).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|4|5] [--precision tf32|fp32]

One "step" = one pass of the hot path over one batch of synthetic 16 kHz audio.  1 frame = one 10 ms hop (160 samples).
--config selects the BASELINE.json configuration (default 2 = configs[1], the one the metric is quoted on):

  2  ConformerCTC(S) offline greedy, 32 x 10 s per GPU (weak scaling)            wav -> mel -> encoder -> CTC decoder -> greedy ids
  3  StreamingConformerCTC block streaming, 64 x 30 s in total, utterances sharded over the N GPUs (strong scaling)
  4  ChunkConformer causal chunk streaming with state caches, 64 streams per GPU, 320 ms chunks (weak scaling; random-init weights:
     the reference ships none)                                                  one step = picker step + feature_pick + decoder step
  5  ConformerCTC(S) + prefix beam search (beam 16), 128 x 5 s in total, sharded over the N GPUs (strong scaling)

`value` = frames/s with the waveforms resident in HBM (device-timed with CUDA events, max over ranks); `e2e` = the same through the
host-buffer call (pinned H2D of every step's input and D2H of its result inside the timed region); `sustained` = the device-resident
loop repeated for >= 2 s.  N > 1 (torchrun): the only collective is ONE async all_gather of the decoded ids + lengths per step
(tensorflowasr_b200/sharding.py), overlapped with the next step.
`--impl reference` times the reference's own CPU deployment path (its shipped ONNX graphs through its vendored onnxruntime 1.10.0,
oracle/_ref; config 4: the NumPy port of the reference model, there being neither weights nor graphs) on the same configuration.
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

# name, batch (per GPU if weak else global), seconds per utterance, scaling, seed, every n-th row is tiled speech
CONFIGS = {
    2: dict(name="ConformerCTC(S) 10M offline greedy, batch 32 x 10 s synthetic 16 kHz per GPU (BASELINE.json configs[1])", kind="offline",
            B=32, seconds=10, scaling="weak", seed=1234, speech_every=4),
    3: dict(name="StreamingConformerCTC 15M block streaming (8000-sample chunks), batch 64 x 30 s in total, utterance shard over the GPUs "
                 "(BASELINE.json configs[2])", kind="streaming", B=64, seconds=30, scaling="strong", seed=1235, speech_every=16),
    4: dict(name="ChunkConformer 15M causal chunk streaming with state caches, 320 ms chunks (chunk_num 32), 64 streams per GPU "
                 "(BASELINE.json configs[3]); random-init weights (the reference ships none)", kind="chunk", B=64, seconds=30, scaling="weak",
            seed=1236, speech_every=4),
    5: dict(name="ConformerCTC(S) + prefix beam search (beam 16, cutoff_top_n 40), batch 128 x 5 s in total, utterance shard over the GPUs "
                 "(BASELINE.json configs[4])", kind="beam", B=128, seconds=5, scaling="strong", seed=1237, speech_every=8),
}


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


def model_dir(kind: str = "offline"):
    """Directory with the reference's encoder.onnx / ctc_model.onnx (the trained weights the engine imports) -- resolved here, without
    touching oracle/: $B200ASR_MODEL_ROOT, the staged copy that travels to the GPU box, or the reference checkout."""
    for base in (os.environ.get("B200ASR_MODEL_ROOT", ""), os.path.join(ROOT, "oracle", "_ref", "models"),
                 "/root/reference/Inference/PythonInference/asr/models"):
        if base and os.path.isfile(os.path.join(base, kind, "encoder.onnx")):
            return os.path.join(base, kind)
    return None


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
        pw = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def algorithmic_gflop(kind: str, B: int, seconds: int) -> float:
    """SURVEY 8(d) algorithmic work (2 x MAC, FFT instead of the dense DFT) of one step of `B` utterances / streams."""
    if kind == "offline":
        return 6.624 * seconds / 10.0 * B
    if kind == "beam":
        return 4.237 * seconds / 5.0 * B
    if kind == "streaming":
        return 41.9 * seconds / 30.0 * B
    # chunk (320 ms step, per stream): 8 encoder frames x 16 blocks x 0.6156 MMAC (= 153.9 M / 250 frames) + subsampler 8 x (3.73 + 0.41) MMAC
    # + helper / decoder blocks and the 9171-class head on <= 16 frames (all frames picked): ~153 MMAC = 0.306 GFLOP
    return 0.306 * B


# ----------------------------------------------------------------------------------------------------------- CPU reference legs
def _ort_threads_best(make, probe):
    """ONNX Runtime 1.10 loses throughput when its intra-op pool is oversubscribed (128 threads ran 10x slower than 16 on the GPU box's
    host): one timed pass per candidate thread count, keep the fastest.  Returns (object, threads)."""
    cores = os.cpu_count() or 1
    best = None
    for th in sorted({min(cores, c) for c in (8, 16, 32, 64, cores)}):
        obj = make(th)
        probe(obj, True)
        t0 = time.perf_counter()
        probe(obj, False)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, obj, th)
    return best[1], best[2]


def cpu_reference_step(config: int):
    """-> (step() -> None, frames per step, cores, kind, description): the reference's own CPU implementation of one step of `config`
    (a bounded sample where a whole step would take minutes; the description says which)."""
    from oracle import ort_ref, ctc_ref
    c = CONFIGS[config]
    Ls = c["seconds"] * SR
    if c["kind"] in ("offline", "beam"):
        if not ort_ref.available():
            return None
        x = synth_batch(c["seed"], c["B"], Ls, c["speech_every"])
        probe_x = x[:4]
        ref, cores = _ort_threads_best(lambda th: ort_ref.ReferenceASR("offline", threads=th),
                                       lambda r, warm: r.logits(r.encode(probe_x[:1] if warm else probe_x)))
        if c["kind"] == "offline":
            def step():
                logits = ref.logits(ref.encode(x))
                return [ctc_ref.greedy_decode(l, logits.shape[-1] - 1) for l in logits]
            return step, c["B"] * Ls // HOP, cores, "reference", (f"the full {c['B']} x {c['seconds']} s batch per step: ONNX Runtime 1.10.0 "
                                                                   f"({cores} of {os.cpu_count()} host threads, the fastest of 8/16/32/64/all) + host greedy decode")
        from concurrent.futures import ThreadPoolExecutor
        from oracle import ctcdec_ref
        pool = ThreadPoolExecutor(max_workers=min(cores, 32))

        def step():
            logits = ref.logits(ref.encode(x))
            probs = ctc_ref.softmax(logits.astype(np.float32)).astype(np.float64)
            return list(pool.map(lambda p: ctcdec_ref.beam_search(p, 16), probs))   # ctc_beam_search_decoder_batch: a thread pool over utterances
        return step, c["B"] * Ls // HOP, cores, "reference", (f"the full {c['B']} x {c['seconds']} s batch per step: ONNX Runtime 1.10.0 ({cores} threads) + the "
                                                               "reference's C++ ctc_beam_search_decoder (beam 16) on a pool of host threads")
    if c["kind"] == "streaming":
        sd = ort_ref.model_dir("streaming")
        if not ort_ref.available() or sd is None:
            return None
        nb = 8                                            # bounded sample: 8 of the 64 utterances (a full step is ~2.7 TFLOP on the CPU)
        x = synth_batch(c["seed"], c["B"], Ls, c["speech_every"])[:nb]
        chunks = x.reshape(nb * (Ls // 8000), 8000, 1)

        def make(th):
            return (ort_ref.OrtModel(os.path.join(sd, "encoder.onnx"), th), ort_ref.OrtModel(os.path.join(sd, "ctc_model.onnx"), th))

        def probe(m, warm):
            e = m[0].run({"inputs": chunks[:4] if warm else chunks[:60]})
            m[1].run({"inputs": e.reshape(1, -1, e.shape[-1])})
        m, cores = _ort_threads_best(make, probe)

        def step():
            enc = m[0].run({"inputs": chunks})
            logits = m[1].run({"inputs": enc.reshape(nb, -1, enc.shape[-1])})
            return [ctc_ref.greedy_decode(l, logits.shape[-1] - 1) for l in logits]
        return step, nb * Ls // HOP, cores, "reference", (f"{nb} of the {c['B']} utterances per step (30 s each = 60 chunks): the reference's streaming ONNX "
                                                          f"graphs through ONNX Runtime 1.10.0 ({cores} threads) + host greedy decode")
    # chunk model: no weights, no graphs -> the NumPy port of the reference model (float64), a few streams
    from oracle import chunk_conformer_ref as cc
    from tensorflowasr_b200 import weights as W
    _, fe_raw, _, _ = W.random_model(0, num_blocks=1)
    cfg = dict(cc.CFG, chunk_num=32)
    raw = cc.random_chunk_model(0, fe_raw, cfg)
    nb = 4
    S = 32 * HOP
    wav = synth_batch(c["seed"], nb, 8 * S, c["speech_every"]).astype(np.float64)
    state = {"c1": cc.init_picker_caches(nb, cfg), "c2": cc.init_decoder_caches(nb, cfg), "i": 0}

    def step():
        i = state["i"] % 8
        state["i"] += 1
        ph, _, hid, state["c1"] = cc.picker_stream_predict(wav[:, i * S:(i + 1) * S], state["c1"], raw, cfg)
        feats, _ = cc.feature_pick(hid, ph, cfg["phone_classes"] - 1)
        if feats.shape[1]:
            _, _, state["c2"] = cc.decoder_stream_predict(feats, state["c2"], raw, cfg)
    return step, nb * 32, 1, "port", f"{nb} streams, one 320 ms step each: NumPy float64 port of the reference model (oracle/chunk_conformer_ref.py)"


def reference_arm(args, rank: int):
    """`--impl reference`: the reference's own CPU path on this configuration (rank 0 only)."""
    if rank != 0:
        return
    c = CONFIGS[args.config]
    line = {"impl": "reference", "metric": METRIC, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gpu_launches": 0}
    got = cpu_reference_step(args.config)
    if got is None:
        line["unavailable"] = "oracle/_ref (vendored onnxruntime + ONNX graphs) was not staged in this checkout"
        print(json.dumps(line))
        return
    step, frames, cores, kind, desc = got
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    n = 0
    while n < args.steps and (n < 2 or time.perf_counter() - t0 < 150.0):     # keep the whole run within a few minutes
        step()
        n += 1
    dt = time.perf_counter() - t0
    val = frames * n / dt
    line.update({"value": val, "ms_per_step": dt / n * 1e3, "steps": n, "rtf": dt / n / (frames * HOP / SR),
                 "config": {"workload": c["name"], "global_batch": c["B"], "seq_len": c["seconds"] * SR, "reference_step": desc},
                 "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": kind, "sample": f"{n} steps; {desc}"},
                 "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def cpu_baseline_sample(config: int):
    got = cpu_reference_step(config)
    if got is None:
        return None
    step, frames, cores, kind, desc = got
    step()
    n, t0 = 0, time.perf_counter()
    while n < 2 or (time.perf_counter() - t0 < 12.0 and n < 40):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": frames * n / dt, "unit": "frames/s", "cores": cores, "kind": kind, "sample": f"{n} steps in {dt:.1f} s; {desc}"}


# ----------------------------------------------------------------------------------------------------------- own arm
def init_dist(local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    if world <= 1:
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # NCCL prints its version (and, with NCCL_DEBUG=INFO, much more) on stdout while the communicator comes up; stdout must carry only
    # the JSON line, so file descriptor 1 points at stderr until the first collective has completed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        import datetime
        # a collective that never completes (a rank died, a mismatch) must fail within minutes, not after the 10-minute default
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)


class Timer:
    """Warm-up, K device-timed steps (CUDA events on the launching stream, barrier + synchronize on both sides), then the same loop
    repeated for >= `sustain` seconds."""

    def __init__(self, world: int, args):
        self.world, self.args = world, args
        self.fork = self.join = lambda: None      # side streams (two batches in flight) wait for / are waited by the timing stream

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_timed(self, step, warm: int, launch_count):
        import torch
        for i in range(warm):
            step(i)
        self.barrier()
        l0 = launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.fork()
        for i in range(self.args.steps):
            step(i)
        self.join()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        launches = launch_count() - l0
        # sustained: repeat for >= sustain seconds (chunks of K steps so that the host never runs more than K steps ahead)
        sus = None
        if self.args.sustain > 0:
            # the number of chunks must be the SAME on every rank (each step issues a collective): derive it from the max-reduced burst time
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            chunks = int(min(max(1.0, -(-self.args.sustain * 1e3 // max(float(t[0]), 1e-3))), 5000))
            n, total_ms = 0, 0.0
            for _ in range(chunks):
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                self.fork()
                for i in range(self.args.steps):
                    step(n + i)
                self.join()
                s1.record()
                s1.synchronize()
                total_ms += s0.elapsed_time(s1)
                n += self.args.steps
            self.barrier()
            sus = (total_ms, n)
        return ms, launches, sus

    def wall_timed(self, step, warm: int):
        for i in range(warm):
            step(i)
        self.barrier()
        t0 = time.perf_counter()
        for i in range(self.args.steps):
            step(i)
        self.barrier()
        return (time.perf_counter() - t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the sustained loop (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=2, choices=[1, 2, 3, 4],
                    help="batches in flight (configs 2/3; sharding.InFlight): n engine handles on n streams, step i on handle i %% n; the single-"
                         "batch-in-flight time is reported beside it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from tensorflowasr_b200 import engine as E, sharding, weights as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    init_dist(local_rank, world)
    cfg = CONFIGS[args.config]
    kind = cfg["kind"]
    Ls = cfg["seconds"] * SR
    if cfg["scaling"] == "weak":
        B_local, B_global, row0 = cfg["B"], cfg["B"] * world, 0
    else:
        b0, b1 = sharding.shard_range(cfg["B"], rank, world)
        B_local, B_global, row0 = b1 - b0, cfg["B"], b0
        if cfg["B"] % world:
            raise SystemExit(f"config {args.config}: {cfg['B']} utterances do not split evenly over {world} GPUs")
    prec = E.PRECISION_TF32 if args.precision == "tf32" else E.PRECISION_FP32
    timer = Timer(world, args)
    stream = torch.cuda.Stream()
    NROT = 8 if kind in ("offline", "beam") else 4
    extra = {}

    # ------------------------------------------------------------------------------------------------ engines + inputs
    if kind == "chunk":
        from tensorflowasr_b200 import chunk_model as CM
        _, fe_raw, _, _ = W.random_model(0, num_blocks=1)
        geo = CM.ChunkGeometry(chunk_num=32)
        eng = CM.ChunkEngine(CM.random_chunk_weights(0, fe_raw, geo), geo, device=local_rank)
        runner = CM.ChunkConformer(eng)
        weights_desc = "random-init ChunkConformer(S) architecture (asr/configs/chunk_conformerS.yml; the reference ships no weights)"
        S = geo.samples_per_step
        nsteps_audio = Ls // S
        host = [torch.from_numpy(synth_batch(cfg["seed"] + 97 * rank + i, B_local, nsteps_audio * S, cfg["speech_every"])).pin_memory() for i in range(2)]
        dev = [h.cuda(non_blocking=True) for h in host]
        frames_per_step_local = B_local * geo.chunk_num
    else:
        md = model_dir("streaming" if kind == "streaming" else "offline")
        if md is not None:
            ge, re_ = W.import_encoder(os.path.join(md, "encoder.onnx"))
            gc, rc = W.import_ctc_model(os.path.join(md, "ctc_model.onnx"))
            weights_desc = "reference-trained weights (shipped ONNX graphs)"
        else:
            if kind == "streaming":
                ge, re_, gc, rc = W.random_model(0, dmodel=256, num_blocks=4, num_heads=4, head_size=64, kernel_size=5, vocab=1332)
            else:
                ge, re_, gc, rc = W.random_model(0, num_blocks=13)
            weights_desc = "random-init weights of the architecture"
        eng = E.Engine(ge, re_, gc, rc, device=local_rank, precision=prec, use_cuda_graph=True,
                       chunk_samples=8000 if kind == "streaming" else 0)
        eng.reserve(B_local, Ls)
        Tp = eng.out_frames(Ls)
        pool = None
        engs = [eng]
        if args.inflight > 1 and kind in ("offline", "streaming"):
            # several batches in flight (sharding.InFlight): one handle per batch in flight (own workspace + graph cache, same weights), each on
            # its own stream; the kernels of step i+1 fill the SMs and the launch gaps step i leaves idle
            made = [eng]

            def make_engine():
                if made:
                    return made.pop()
                e2 = E.Engine(ge, re_, gc, rc, device=local_rank, precision=prec, use_cuda_graph=True, chunk_samples=8000 if kind == "streaming" else 0)
                e2.reserve(B_local, Ls)
                return e2
            pool = sharding.InFlight(make_engine, args.inflight, torch.device("cuda", local_rank))
            engs = pool.engines
            timer.fork, timer.join = pool.fork, pool.join
        # inputs: NROT distinct batches (> L2 together) rotated so that no step re-reads a cached waveform
        host = [torch.from_numpy(synth_batch(cfg["seed"] + 97 * rank + i, cfg["B"] if cfg["scaling"] == "strong" else B_local, Ls,
                                             cfg["speech_every"])[row0:row0 + B_local]).pin_memory() for i in range(NROT)]
        dev = [h.cuda(non_blocking=True) for h in host]
        frames_per_step_local = B_local * (Ls // HOP)

    with torch.cuda.stream(stream):
        # -------------------------------------------------------------------------------------------- step functions
        if kind in ("offline", "streaming"):
            xch = sharding.IdsExchange(B_local, Tp, torch.device("cuda", local_rank))
            xchs = [xch] + [sharding.IdsExchange(B_local, Tp, torch.device("cuda", local_rank)) for _ in engs[1:]]
            NI = len(engs)

            def step1(i):                     # one batch in flight: everything on the timing stream
                sl = xch.acquire()
                ids, lens = xch.buffers(sl)
                eng.recognize(dev[i % NROT], ids, lens)
                xch.gather(sl)
                return sl

            def stepn(i):                     # batch i on handle / stream i % NI
                k = i % NI
                with torch.cuda.stream(pool.stream_of(i)):
                    sl = xchs[k].acquire()
                    ids, lens = xchs[k].buffers(sl)
                    engs[k].recognize(dev[i % NROT], ids, lens)
                    xchs[k].gather(sl)
                return sl
            step = stepn if pool is not None else step1

            hres = [torch.empty((world * B_local * (Tp + 1),), dtype=torch.int32).pin_memory() for _ in range(2)]
            if world == 1:

                hid2 = [torch.empty((B_local, Tp), dtype=torch.int32).pin_memory() for _ in range(2 * NI)]
                hlen2 = [torch.empty((B_local,), dtype=torch.int32).pin_memory() for _ in range(2 * NI)]

                def e2e_run(n, base):       # two-deep pipeline of the C ABI per handle: H2D of step i+1 under the compute of step i
                    D = 2 * NI
                    for i in range(n):
                        k, sl = i % NI, (i // NI) & 1
                        if i >= D:
                            engs[k].recognize_host_collect(sl)
                        engs[k].recognize_host_submit(sl, host[(base + i) % NROT], hid2[i % D], hlen2[i % D])
                    for i in range(max(n - D, 0), n):
                        engs[i % NI].recognize_host_collect((i // NI) & 1)
                e2e_mode = ("two-deep pipeline (b200asr_recognize_host_submit/_collect): H2D of step i+1 under the compute of step i"
                            + (f", alternating over {NI} handles" if NI > 1 else ""))
            else:
                copy_stream = torch.cuda.Stream()
                dev_in2 = [torch.empty_like(dev[0]) for _ in range(2)]
                h2d_done = [torch.cuda.Event() for _ in range(2)]
                consumed = [torch.cuda.Event() for _ in range(2)]
                done = [torch.cuda.Event() for _ in range(2)]
                for ev in consumed:
                    ev.record(stream)

                def e2e_run(n, base):       # the same two-deep overlap from the Python API + the async ids exchange
                    for i in range(n):
                        sl = i & 1
                        if i >= 2:
                            done[sl].synchronize()
                        with torch.cuda.stream(copy_stream):
                            copy_stream.wait_event(consumed[sl])
                            dev_in2[sl].copy_(host[(base + i) % NROT], non_blocking=True)
                            h2d_done[sl].record(copy_stream)
                        stream.wait_event(h2d_done[sl])
                        xs = xch.acquire()
                        ids, lens = xch.buffers(xs)
                        eng.recognize(dev_in2[sl], ids, lens)
                        consumed[sl].record(stream)
                        xch.gather(xs)
                        xch.result(xs)                                    # (stream waits for the collective: the D2H below reads it)
                        hres[sl].copy_(xch.gathered[xs], non_blocking=True)
                        done[sl].record(stream)
                    for i in range(max(n - 2, 0), n):
                        done[i & 1].synchronize()
                e2e_mode = "two-deep pipeline (copy stream + Engine.recognize + one async all_gather of ids+lengths): H2D of step i+1 under step i"
            h2d_bytes, d2h_bytes = B_local * Ls * 4, world * B_local * (Tp + 1) * 4
            launch_count = lambda: sum(e_.launch_count for e_ in engs)
        elif kind == "beam":
            BEAM = 16
            enc = torch.empty((B_local, Tp, ge.dmodel), device="cuda", dtype=torch.float32)
            logits = torch.empty((B_local, Tp, gc.vocab), device="cuda", dtype=torch.float32)
            xch = sharding.IdsExchange(B_local, Tp, torch.device("cuda", local_rank))
            dev_in = torch.empty_like(dev[0])
            hres = torch.empty((world * B_local * (Tp + 1),), dtype=torch.int32).pin_memory()

            def decode(x):
                eng.encode(x, out=enc)
                eng.ctc_logits(enc, out=logits)
                ids, lens, scores = eng.ctc_beam(logits, BEAM)
                sl = xch.acquire()
                bi, bl = xch.buffers(sl)
                bi.copy_(ids[:, 0])                                       # the best hypothesis of every utterance is what is exchanged
                bl.copy_(lens[:, 0])
                xch.gather(sl)
                return sl

            def step(i):
                return decode(dev[i % NROT])

            def e2e_run(n, base):
                for i in range(n):
                    dev_in.copy_(host[(base + i) % NROT], non_blocking=True)
                    sl = decode(dev_in)
                    xch.result(sl)
                    hres.copy_(xch.gathered[sl], non_blocking=True)
                    torch.cuda.current_stream().synchronize()
            e2e_mode = "synchronous per step: H2D, encode, CTC logits, beam search, ids exchange, D2H of the best hypotheses"
            h2d_bytes, d2h_bytes = B_local * Ls * 4, world * B_local * (Tp + 1) * 4
            launch_count = lambda: eng.launch_count
        else:  # chunk streaming
            T = geo.frames_per_step
            state = runner.init_picker_caches(B_local)
            tok = torch.full((B_local, geo.dec_back + 4 * T + 1), -1, device="cuda", dtype=torch.int32)
            xch = sharding.IdsExchange(B_local, geo.dec_back + 4 * T, torch.device("cuda", local_rank))
            stats = {"dec_steps": 0, "picked": 0}
            hchunks = [host[0][:, k * S:(k + 1) * S].contiguous().pin_memory() for k in range(nsteps_audio)]   # one pinned buffer per 320 ms step
            dchunk = torch.empty((B_local, S), device="cuda")
            hres = torch.empty((world * B_local * (geo.dec_back + 4 * T + 1),), dtype=torch.int32).pin_memory()

            def one(chunk):
                ph, _, hid, _ = runner.picker_stream_predict(chunk, state)
                feats, _ = runner.feature_pick(hid, ph)                   # (synchronises: the decoder's shape is data dependent)
                sl = xch.acquire()
                ids, lens = xch.buffers(sl)
                ids.fill_(-1)
                lens.zero_()
                if feats.shape[1]:
                    valid, unvalid, _ = runner.decoder_stream_predict(feats, state)
                    nv = valid.shape[1]
                    if nv:
                        ids[:, :nv] = valid.argmax(-1).to(torch.int32)
                        lens.fill_(nv)
                    stats["dec_steps"] += 1
                    stats["picked"] += feats.shape[1]
                xch.gather(sl)
                return sl

            def step(i):
                k = i % nsteps_audio
                return one(dev[(i // nsteps_audio) % 2][:, k * S:(k + 1) * S])

            def e2e_run(n, base):
                for i in range(n):
                    k = (base + i) % nsteps_audio
                    dchunk.copy_(hchunks[k], non_blocking=True)
                    sl = one(dchunk)
                    xch.result(sl)
                    hres.copy_(xch.gathered[sl], non_blocking=True)
                    torch.cuda.current_stream().synchronize()
            e2e_mode = "per step: H2D of the 320 ms chunk of every stream, picker step, feature_pick, decoder step, ids exchange, D2H of the token ids"
            h2d_bytes, d2h_bytes = B_local * S * 4, world * B_local * (geo.dec_back + 4 * T + 1) * 4
            launch_count = lambda: eng.launch_count

        # -------------------------------------------------------------------------------------------- timing
        warm = max(args.warmup, NROT) if kind != "chunk" else max(args.warmup, 24)    # every rotated input / cache-fill state has its own captured graph
        if kind in ("offline", "streaming") and pool is not None:
            import math
            warm = max(warm, NROT * len(pool) // math.gcd(NROT, len(pool)))          # ... per handle: every (handle, input) pair once
        sampler = ClockSampler(local_rank)
        for i in range(warm):
            step(i)
        timer.barrier()
        sampler.start()
        ms_total, launches, sus = timer.device_timed(step, 0, launch_count)
        if kind in ("offline", "streaming") and pool is not None:
            # the same K steps with ONE batch in flight (handle 0 alone, on the timing stream): the per-batch latency view of the step
            for i in range(NROT):
                step1(i)
            timer.barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(args.steps):
                step1(i)
            s1.record()
            timer.barrier()
            extra["single_batch_in_flight"] = {"ms_per_step": s0.elapsed_time(s1) / args.steps,
                                               "note": "same loop on one handle / one stream: what one batch costs when nothing else shares the GPU"}
        e2e_run(max(args.warmup, 2 * (len(pool) if kind in ("offline", "streaming") and pool is not None else 1)), 0)   # every (handle, pipeline slot) once: its graph is captured here
        timer.barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps, 3)
        timer.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        sync_ms = None
        if kind in ("offline", "streaming") and world == 1:
            hid = torch.empty((B_local, Tp), dtype=torch.int32).pin_memory()
            hlen = torch.empty((B_local,), dtype=torch.int32).pin_memory()
            sync_ms = timer.wall_timed(lambda i: eng.recognize_host(host[i % NROT], hid, hlen), args.warmup)
        sampler.stop_flag.set()
        sampler.join(timeout=2)

        # -------------------------------------------------------------------------------------------- roofline (rank 0)
        roof = None
        if rank == 0:
            hbm_peak, tf_peak, tf_sus, how = measured_peaks()
            step_ms = ms_total / args.steps
            gflop = algorithmic_gflop(kind, B_local, cfg["seconds"])
            whole = {"algorithmic_gflop_per_step": gflop, "tflops": gflop / step_ms, "frac_of_bf16_burst_peak": gflop / step_ms / tf_peak,
                     "note": "SURVEY 8(d) algorithmic work (FFT, not the dense DFT) / device-timed step"}
            if sus:
                whole["sustained_tflops"] = gflop / (sus[0] / sus[1])
                whole["frac_of_bf16_sustained_peak"] = gflop / (sus[0] / sus[1]) / tf_sus
            if kind != "chunk":
                traffic = {}
                tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram bytes per launch from the committed ncu --set full capture
                if os.path.isfile(tpath):
                    traffic = json.load(open(tpath))
                eng.recognize(dev[0])            # (no collective here: only rank 0 runs the roofline section) leaves the last block's operands in the workspace
                torch.cuda.synchronize()

                def stage_line(st, iters=20):
                    m, f, b = eng.time_stage(st, B_local, Ls, iters=iters)
                    return m, f, b, {"ms_per_launch": round(m, 5), "tflops": round(f / (m * 1e-3) / 1e12, 2),
                                     "frac_tensor_peak": round(f / (m * 1e-3) / 1e12 / tf_peak, 4), "gbs": round(b / (m * 1e-3) / 1e9, 1),
                                     "frac_hbm_peak": round(b / (m * 1e-3) / 1e9 / hbm_peak, 4), "flops_per_launch": f,
                                     "algorithmic_bytes_per_launch": b, "traffic": traffic.get(st) if args.config == 2 else None}

                ms_k, flops, bytes_, _ = stage_line("ffn_chain")
                ach = flops / (ms_k * 1e-3) / 1e12
                nblk = ge.num_blocks + gc.num_blocks
                roof = {"kernel": f"FFModule as one chained tcgen05 kernel (both GEMMs + swish + 0.5-residual + LayerNorm), M = {eng.out_frames(Ls) * B_local} rows, "
                                  f"d_model {ge.dmodel}: {2 * nblk} launches of a step",
                        "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                        "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({how}); the kernel computes in tf32 (half the bf16 tensor rate: frac 0.5 = the tf32 ceiling)",
                        "ms_per_launch": ms_k, "flops_per_launch": flops, "algorithmic_bytes_per_launch": bytes_,
                        "traffic": traffic.get("ffn_chain") if args.config == 2 else None, "whole_step": whole}
                others = {}
                for st in ("conv2", "stft", "conv1", "qkv", "attention", "dwconv", "sub_linear", "ctc_fc"):
                    try:
                        others[st] = stage_line(st)[3]
                    except RuntimeError as ex:            # e.g. conv1 has no kernel of its own when it is fused into conv2's
                        others[st] = {"skipped": str(ex)[:160]}
                roof["other_stages"] = others
            else:
                roof = {"kernel": "whole streaming step (picker step + feature_pick + decoder step): ~330 launches on B x 8 = 512 rows -- launch-latency "
                                  "bound, no single dominant kernel", "bound": "tensor", "achieved": whole["tflops"], "peak": tf_peak, "unit": "TFLOP/s",
                        "frac": whole["frac_of_bf16_burst_peak"], "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({how})", "traffic": None, "whole_step": whole}
                extra["stream_stats"] = {"decoder_steps_in_timed_and_warm_loops": stats["dec_steps"], "frames_picked": stats["picked"]}

    t = torch.tensor([ms_total, e2e_ms, sus[0] if sus else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, sus_ms = float(t[0]), float(t[1]), float(t[2])
    frames_step = frames_per_step_local * world
    if rank == 0:
        value = frames_step * args.steps / (ms_total * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
                "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
                "rtf": (ms_total * 1e-3 / args.steps) / (frames_step * HOP / SR),
                "config": {"workload": cfg["name"], "baseline_config": args.config, "weights": weights_desc, "global_batch": B_global,
                           "seq_len": Ls, "parallelism": f"dp{world} (utterance shard, one async all_gather of ids+lengths per step)",
                           "l2_policy": (f"{NROT} distinct input batches rotated ({NROT * B_local * Ls * 4 / 1e6:.0f} MB > 126 MB L2)" if kind != "chunk" else
                                         "every step reads a new 320 ms chunk per stream; caches + weights (~140 MB) exceed L2"),
                           "frame": "10 ms hop (160 samples)", "batches_in_flight": (args.inflight if kind in ("offline", "streaming") else 1),
                           "operands": ("tf32 tensor-core operands, fp32 accumulate and fp32 activations in HBM; the subsampler's conv1 map and conv2 output are "
                                        "stored as IEEE fp16 (the same 11-bit significand as tf32) and its two GEMMs run kind::f16" if args.precision == "tf32" and kind != "chunk"
                                        else ("tf32 tensor-core operands, fp32 accumulate" if args.precision == "tf32" else "fp32 CUDA cores"))},
                "e2e": {"value": frames_step * args.steps / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms / args.steps, "mode": e2e_mode},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof}
        if sync_ms is not None:
            line["e2e"]["sync_call_ms_per_step"] = sync_ms / args.steps
        if sus:
            line["sustained"] = {"seconds": sus_ms * 1e-3, "steps": sus[1], "ms_per_step": sus_ms / sus[1],
                                 "value": frames_step * sus[1] / (sus_ms * 1e-3), "unit": "frames/s",
                                 "note": "the device-resident loop repeated for >= 2 s (clocks / power settle); compare with bf16_tflops_sustained"}
        if kind == "chunk":
            line["latency_ms_per_stream_step"] = ms_total / args.steps
        line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_sample(args.config)
            except Exception as ex:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"error": str(ex)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
