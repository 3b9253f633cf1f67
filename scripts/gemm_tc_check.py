"""tcgen05 GEMM vs fp32 CUDA-core GEMM vs torch fp64 on the GPU (run under gpurun with a timeout)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorflowasr_b200 import engine as E, weights as W

ge, re_, gc, rc = W.random_model(0, num_blocks=1)
eng = E.Engine(ge, re_, gc, rc, precision=0, use_cuda_graph=False)
torch.manual_seed(0)
shapes = [(128, 144, 144), (130, 144, 32), (8000, 576, 144), (8000, 144, 576), (100, 144, 2880), (8000, 1332, 144), (300, 256, 256),
          (1000, 432, 144), (77, 64, 64), (8000, 288, 144), (513, 1024, 256), (4097, 128, 96), (1, 144, 144)]
ok = True
for (M, N, K) in shapes:
    A = torch.randn(M, K, device="cuda")
    Wt = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    ref64 = A.double() @ Wt.double().T
    for epi in (0, 1, 2, 3, 4, 5):
        if epi == 3 and N % 8:
            continue
        r = ref64 + (bias.double() if epi != 5 else 0)
        if epi == 1: r = r.clamp_min(0)
        if epi == 2: r = r * torch.sigmoid(r)
        if epi == 3: r = r[:, 0::2] * torch.sigmoid(r[:, 1::2])
        if epi == 4: r = resid.double() + 0.5 * r
        c_tc = eng.debug_gemm(A, Wt, bias, resid, 0.5, epi, True)
        c_32 = eng.debug_gemm(A, Wt, bias, resid, 0.5, epi, False)
        torch.cuda.synchronize()
        e_tc = (c_tc.double() - r).abs().max().item()
        e_32 = (c_32.double() - r).abs().max().item()
        bad = (e_tc > 2e-2) or (e_32 > 1e-4) or torch.isnan(c_tc).any().item()
        ok &= not bad
        print(f"M={M:5d} N={N:5d} K={K:5d} epi={epi} err_tc={e_tc:.3e} err_fp32={e_32:.3e} mean_tc_signed={(c_tc.double()-r).mean().item():+.2e} {'BAD' if bad else ''}", flush=True)
# fused LayerNorm epilogues
def ln(x, g, b, eps=1e-3):
    mu = x.mean(-1, keepdim=True); var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b
for (M, N, K) in [(8000, 144, 576), (300, 144, 144), (1000, 256, 1024), (129, 64, 128), (8000, 144, 2880)]:
    A = torch.randn(M, K, device="cuda"); Wt = torch.randn(N, K, device="cuda") / K ** 0.5; bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda") * 30
    g1, b1, g2, b2 = (torch.randn(N, device="cuda") for _ in range(4))
    acc = A.double() @ Wt.double().T + bias.double()
    for epi, inplace in ((6, False), (6, True), (7, False), (7, True), (8, False)):
        r = resid.clone()
        x = (r.double() + 0.5 * acc) if epi != 8 else acc
        if epi == 7:
            c_ref = ln(x, g1.double(), b1.double()); c2_ref = ln(c_ref, g2.double(), b2.double())
        else:
            c_ref = x; c2_ref = ln(x, g1.double(), b1.double())
        C, C2 = eng.debug_gemm_ln(A, Wt, bias, r if epi != 8 else None, 0.5, epi, (g1, b1), (g2, b2) if epi == 7 else None, inplace=inplace)
        torch.cuda.synchronize()
        e1 = (C.double() - c_ref).abs().max().item(); e2 = (C2.double() - c2_ref).abs().max().item()
        bad = e1 > 3e-2 or e2 > 3e-2 or torch.isnan(C2).any().item()
        ok &= not bad
        print(f"LN M={M} N={N} K={K} epi={epi} inplace={inplace} errC={e1:.3e} errC2={e2:.3e} {'BAD' if bad else ''}", flush=True)
# timing of the big ones
for (M, N, K) in [(8000, 576, 144), (8000, 144, 576), (8000, 1332, 144), (160000, 144, 1296)]:
    A = torch.randn(M, K, device="cuda"); Wt = torch.randn(N, K, device="cuda"); bias = torch.randn(N, device="cuda")
    for tc in (True, False):
        for _ in range(3): eng.debug_gemm(A, Wt, bias, None, 1.0, 0, tc)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(20): eng.debug_gemm(A, Wt, bias, None, 1.0, 0, tc)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
        print(f"time M={M} N={N} K={K} tc={tc}: {dt*1e6:.1f} us  {2*M*N*K/dt/1e12:.1f} TFLOP/s")
print("ALL OK" if ok else "FAILURES")
