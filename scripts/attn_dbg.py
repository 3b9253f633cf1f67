import os, sys
os.environ["B200ASR_ATTN_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorflowasr_b200 import engine as E, weights as W
ge, re_, gc, rc = W.random_model(0, num_blocks=1)
eng = E.Engine(ge, re_, gc, rc, precision=0, use_cuda_graph=False)
B, T, H, dh = 32, 250, 4, 36
qkv = torch.randn(B * T, 3 * H * dh, device="cuda")
for _ in range(3):
    eng.debug_attention(qkv, B, T, H, dh, True); torch.cuda.synchronize()
