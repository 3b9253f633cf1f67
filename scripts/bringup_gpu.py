"""GPU bring-up: stage-by-stage comparison of the CUDA path against the NumPy oracle (run under gpurun)."""
import os, sys, time, wave
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorflowasr_b200 import engine as E, weights as W
from oracle import conformer_ref as cr, ort_ref

prec = int(os.environ.get("PREC", "1"))
wav_path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "BAC009S0764W0121.wav")
w = wave.open(wav_path)
x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768
md = ort_ref.model_dir("offline")
ge, re_ = W.import_encoder(os.path.join(md, "encoder.onnx"))
gc, rc = W.import_ctc_model(os.path.join(md, "ctc_model.onnx"))


def cmp(name, a, b):
    a = np.asarray(a, dtype=np.float64).reshape(b.shape)
    d = np.abs(a - b)
    print(f"{name:14s} max|ref|={np.abs(b).max():10.4f} maxabs={d.max():.3e} rel={d.max() / max(np.abs(b).max(), 1e-30):.3e} nan={np.isnan(a).any()}", flush=True)


eng = E.Engine(ge, re_, gc, rc, precision=prec, use_cuda_graph=False)
taps = {}
enc_ref = cr.encoder_forward(x[None], re_, ge.num_blocks, taps=taps)
lg_ref = cr.ctc_forward(enc_ref, rc, 1)
mel = eng.mel(x[None]).cpu().numpy()
cmp("mel", mel, taps["mel"])
enc = eng.encode(x[None])
cmp("enc", enc.cpu().numpy(), enc_ref)
lg = eng.ctc_logits(enc)
cmp("logits", lg.cpu().numpy(), lg_ref)
ids, lens = eng.ctc_greedy(lg)
print("ids ", ids[0, :int(lens[0])].tolist())
ids2, lens2 = eng.recognize(x[None])
print("ids2", ids2[0, :int(lens2[0])].tolist())
print("gold", [669, 82, 103, 78, 247, 56, 71, 573, 386, 82, 30, 213, 496])
hi, hl = eng.recognize_host(torch.from_numpy(x[None]).pin_memory())
print("host", hi[0, :int(hl[0])].tolist())

# batch of 2 equal-length noise utterances vs oracle
rng = np.random.default_rng(7)
xb = np.clip(rng.standard_normal((2, 32000)).astype(np.float32) * 0.1, -1, 1)
cmp("enc noise b2", eng.encode(xb).cpu().numpy(), cr.encoder_forward(xb, re_, ge.num_blocks))

# throughput, B=32 x 10 s
eng2 = E.Engine(ge, re_, gc, rc, precision=prec, use_cuda_graph=True)
B, L = 32, 160000
xs = torch.from_numpy(np.clip(np.random.default_rng(1234).standard_normal((B, L)).astype(np.float32) * 0.1, -1, 1)).cuda()
ids = torch.empty((B, eng2.out_frames(L)), device="cuda", dtype=torch.int32)
lens = torch.empty((B,), device="cuda", dtype=torch.int32)
for _ in range(3):
    eng2.recognize(xs, ids, lens)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
n = 10
for _ in range(n):
    eng2.recognize(xs, ids, lens)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / n
print(f"recognize B=32x10s: {ms:.3f} ms/batch -> {B * 1000 / ms * 1000:.0f} frames/s ; launches/call={eng2.launch_count // 13}")
