import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorflowasr_b200 import engine as E, weights as W
ge, re_, gc, rc = W.random_model(0, num_blocks=1)
eng = E.Engine(ge, re_, gc, rc, precision=0, use_cuda_graph=False)
M, K1, N1, N2 = 8000, 144, 576, 144
X = torch.randn(M, K1, device="cuda"); W1 = torch.randn(N1, K1, device="cuda") / 12; b1 = torch.randn(N1, device="cuda")
W2 = torch.randn(N2, N1, device="cuda") / 24; b2 = torch.randn(N2, device="cuda"); g1 = torch.randn(N2, device="cuda"); be1 = torch.randn(N2, device="cuda")
resid = torch.randn(M, N2, device="cuda")
for _ in range(3):
    eng.debug_chain(X, W1, b1, W2, b2, resid, 0.5, 6, (g1, be1))
    eng.debug_gemm(X, W1, b1, None, 1.0, 2, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for _ in range(20): eng.debug_chain(X, W1, b1, W2, b2, resid, 0.5, 6, (g1, be1))
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("chain kernel (graph of 20):", e0.elapsed_time(e1) / 20 * 1e3, "us each")
