import os, sys
os.environ["B200ASR_PAIR_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorflowasr_b200 import engine as E, weights as W
ge, re_, gc, rc = W.random_model(0, num_blocks=1)
eng = E.Engine(ge, re_, gc, rc, precision=0, use_cuda_graph=False)
M, K1, N2 = 8000, 144, 144
for N1 in (576, 288):
    X = torch.randn(M, K1, device="cuda"); W1 = torch.randn(N1, K1, device="cuda") / 12; b1 = torch.randn(N1, device="cuda")
    W2 = torch.randn(N2, N1, device="cuda") / 24; b2 = torch.randn(N2, device="cuda"); g1 = torch.randn(N2, device="cuda"); be1 = torch.randn(N2, device="cuda")
    resid = torch.randn(M, N2, device="cuda")
    print("N1 =", N1, file=sys.stderr)
    for _ in range(3):
        eng.debug_chain(X, W1, b1, W2, b2, resid, 0.5, 6, (g1, be1), pair=True)
        torch.cuda.synchronize()
