"""Launch the subsampling convolutions alone on the benchmark shape (for ncu / A-B): B200ASR_FUSED_SUB=1 selects the fused kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tensorflowasr_b200 import engine as E, weights as W
md = bench.model_dir("offline")
ge, re_ = W.import_encoder(os.path.join(md, "encoder.onnx"))
gc, rc = W.import_ctc_model(os.path.join(md, "ctc_model.onnx"))
eng = E.Engine(ge, re_, gc, rc, use_cuda_graph=False)
mel = torch.randn(32, 1000, 80, device="cuda") * 20 - 40
for _ in range(3):
    out = eng.debug_subsample_convs(mel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = eng.debug_subsample_convs(mel)
e1.record(); torch.cuda.synchronize()
print("subsample convs: %.1f us per call" % (e0.elapsed_time(e1) * 100))
