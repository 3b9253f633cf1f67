"""tcgen05 attention vs fp32 CUDA-core attention vs torch fp64 (run under gpurun with a timeout)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorflowasr_b200 import engine as E, weights as W
ge, re_, gc, rc = W.random_model(0, num_blocks=1)
eng = E.Engine(ge, re_, gc, rc, precision=0, use_cuda_graph=False)
torch.manual_seed(0)
ok = True
for (B, T, H, dh, wf, wb) in [(2, 250, 4, 36, -1, 0), (1, 106, 4, 36, -1, 0), (3, 13, 4, 64, -1, 0), (1, 300, 2, 32, -1, 0), (2, 780, 4, 64, -1, 0),
                              (1, 1, 4, 36, -1, 0), (2, 257, 4, 36, -1, 0), (2, 120, 4, 36, 36, 0), (1, 300, 4, 36, 36, 8), (32, 250, 4, 36, -1, 0)]:
    qkv = torch.randn(B * T, 3 * H * dh, device="cuda")
    qkv[:, :H * dh] *= 0.5
    q, k, v = (qkv[:, i * H * dh:(i + 1) * H * dh].reshape(B, T, H, dh).double() for i in range(3))
    s = torch.einsum("bnhd,bmhd->bhnm", q, k)
    if wf >= 0:
        i = torch.arange(T, device="cuda")[:, None]; j = torch.arange(T, device="cuda")[None, :]
        lo = torch.clamp(torch.minimum(torch.clamp(i - wf, min=0), torch.tensor(T - wb, device="cuda")), min=0)
        hi = torch.clamp(torch.maximum(torch.minimum(i + wb, torch.tensor(T, device="cuda")), torch.tensor(wb, device="cuda")), max=T - 1)
        s = s.masked_fill(~((j >= lo) & (j <= hi)), float("-inf"))
    ref = torch.einsum("bhnm,bmhd->bnhd", torch.softmax(s, -1), v).reshape(B * T, H * dh)
    o_tc = eng.debug_attention(qkv, B, T, H, dh, True, wf, wb)
    o_32 = eng.debug_attention(qkv, B, T, H, dh, False, wf, wb)
    torch.cuda.synchronize()
    e_tc = (o_tc.double() - ref).abs().max().item(); e_32 = (o_32.double() - ref).abs().max().item()
    bad = e_tc > 1e-2 or e_32 > 1e-4 or torch.isnan(o_tc).any().item()
    ok &= not bad
    print(f"B={B} T={T} H={H} dh={dh} win=({wf},{wb}) err_tc={e_tc:.3e} err_fp32={e_32:.3e} {'BAD' if bad else ''}", flush=True)
B, T, H, dh = 32, 250, 4, 36
qkv = torch.randn(B * T, 3 * H * dh, device="cuda")
for tc in (True, False):
    for _ in range(3): eng.debug_attention(qkv, B, T, H, dh, tc)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): eng.debug_attention(qkv, B, T, H, dh, tc)
    torch.cuda.synchronize(); print(f"time tc={tc}: {(time.perf_counter() - t) / 20 * 1e6:.1f} us")
print("ALL OK" if ok else "FAILURES")
