"""Chained two-GEMM tcgen05 kernel vs torch fp64 (run under gpurun with a timeout)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorflowasr_b200 import engine as E, weights as W
ge, re_, gc, rc = W.random_model(0, num_blocks=1)
eng = E.Engine(ge, re_, gc, rc, precision=0, use_cuda_graph=False)
torch.manual_seed(0)
def ln(x, g, b, eps=1e-3):
    mu = x.mean(-1, keepdim=True); var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b
ok = True
for (M, K1, N1, N2) in [(8000, 144, 576, 144), (8000, 144, 288, 144), (300, 144, 576, 144), (1, 144, 288, 144), (20000, 144, 576, 144),
                        (1000, 256, 1024, 256), (777, 256, 512, 256)]:
    X = torch.randn(M, K1, device="cuda"); W1 = torch.randn(N1, K1, device="cuda") / K1 ** 0.5; b1 = torch.randn(N1, device="cuda") * 0.3
    W2 = torch.randn(N2, N1, device="cuda") / N1 ** 0.5; b2 = torch.randn(N2, device="cuda") * 0.3
    g1, be1, g2, be2 = (torch.randn(N2, device="cuda") for _ in range(4))
    for epi in (6, 7):
        resid = torch.randn(M, N2, device="cuda") * 20
        hid = X.double() @ W1.double().T + b1.double(); hid = hid * torch.sigmoid(hid)
        x = resid.double() + 0.5 * (hid @ W2.double().T + b2.double())
        if epi == 6: c_ref, c2_ref = x, ln(x, g1.double(), be1.double())
        else:
            c_ref = ln(x, g1.double(), be1.double()); c2_ref = ln(c_ref, g2.double(), be2.double())
        C, C2 = eng.debug_chain(X, W1, b1, W2, b2, resid, 0.5, epi, (g1, be1), (g2, be2) if epi == 7 else None)
        torch.cuda.synchronize()
        e1 = (C.double() - c_ref).abs().max().item(); e2 = (C2.double() - c2_ref).abs().max().item()
        bad = e1 > 3e-2 or e2 > 3e-2 or torch.isnan(C2).any().item()
        ok &= not bad
        print(f"chain M={M} K1={K1} N1={N1} N2={N2} epi={epi} errC={e1:.3e} errC2={e2:.3e} {'BAD' if bad else ''}", flush=True)
M, K1, N1, N2 = 8000, 144, 576, 144
X = torch.randn(M, K1, device="cuda"); W1 = torch.randn(N1, K1, device="cuda") / 12; b1 = torch.randn(N1, device="cuda")
W2 = torch.randn(N2, N1, device="cuda") / 24; b2 = torch.randn(N2, device="cuda"); g1 = torch.randn(N2, device="cuda"); be1 = torch.randn(N2, device="cuda")
resid = torch.randn(M, N2, device="cuda")
for _ in range(3): eng.debug_chain(X, W1, b1, W2, b2, resid, 0.5, 6, (g1, be1))
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): eng.debug_chain(X, W1, b1, W2, b2, resid, 0.5, 6, (g1, be1))
torch.cuda.synchronize(); print(f"chain FFN time: {(time.perf_counter() - t) / 20 * 1e6:.1f} us")
print("ALL OK" if ok else "FAILURES")
