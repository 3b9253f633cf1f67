"""GPU kernel-level parity tests (run with -m gpu): every tcgen05 kernel behind the C ABI's test hooks against torch fp64 on the
same seeded inputs, next to the exact-fp32 CUDA-core kernel of the same operation.

Tolerances: tf32 tensor-core path 2e-2 (plain epilogues) / 3e-2 (LayerNorm epilogues, residual scale 30) / 1e-2 (attention);
fp32 CUDA-core path 1e-4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


@pytest.fixture(scope="module")
def keng():
    from tensorflowasr_b200 import engine as E, weights as W
    ge, re_, gc, rc = W.random_model(0, num_blocks=1)
    e = E.Engine(ge, re_, gc, rc, precision=0, use_cuda_graph=False)
    yield e
    e.close()


def _ln(torch, x, g, b, eps=1e-3):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


@pytest.mark.parametrize("M,N,K", [(128, 144, 144), (130, 144, 32), (8000, 576, 144), (8000, 144, 576), (100, 144, 2880), (8000, 1332, 144),
                                   (300, 256, 256), (8000, 432, 144), (1000, 432, 144), (77, 64, 64), (4097, 128, 96), (1, 144, 144)])
def test_tcgen05_gemm_epilogues_vs_fp64(keng, torch_mod, M, N, K):
    """C = epilogue(A W^T): bias / ReLU / swish / GLU / residual / none, tcgen05 tf32 and fp32 CUDA cores (ragged M, N, K tails,
    the two-tile 224-column QKV configuration at M = 8000, N = 432 included)."""
    torch = torch_mod
    torch.manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda")
    Wt = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    ref64 = A.double() @ Wt.double().T
    for epi in (0, 1, 2, 3, 4, 5):
        if epi == 3 and N % 8:
            continue
        r = ref64 + (bias.double() if epi != 5 else 0)
        if epi == 1:
            r = r.clamp_min(0)
        if epi == 2:
            r = r * torch.sigmoid(r)
        if epi == 3:
            r = r[:, 0::2] * torch.sigmoid(r[:, 1::2])
        if epi == 4:
            r = resid.double() + 0.5 * r
        c_tc = keng.debug_gemm(A, Wt, bias, resid, 0.5, epi, True)
        c_32 = keng.debug_gemm(A, Wt, bias, resid, 0.5, epi, False) if K % 16 == 0 else None
        torch.cuda.synchronize()
        assert not torch.isnan(c_tc).any()
        assert (c_tc.double() - r).abs().max().item() < 2e-2, f"tcgen05 epilogue {epi}"
        if c_32 is not None:
            assert (c_32.double() - r).abs().max().item() < 1e-4, f"fp32 epilogue {epi}"


@pytest.mark.parametrize("M,N,K", [(8000, 144, 576), (300, 144, 144), (1000, 256, 1024), (129, 64, 128), (8000, 144, 2880)])
def test_tcgen05_gemm_layernorm_epilogues_vs_fp64(keng, torch_mod, M, N, K):
    """Residual + LayerNorm (one or two, chained) fused in the GEMM epilogue, in place and out of place."""
    torch = torch_mod
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    Wt = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda") * 30
    g1, b1, g2, b2 = (torch.randn(N, device="cuda") for _ in range(4))
    acc = A.double() @ Wt.double().T + bias.double()
    for epi, inplace in ((6, False), (6, True), (7, False), (7, True), (8, False)):
        r = resid.clone()
        x = (r.double() + 0.5 * acc) if epi != 8 else acc
        if epi == 7:
            c_ref = _ln(torch, x, g1.double(), b1.double())
            c2_ref = _ln(torch, c_ref, g2.double(), b2.double())
        else:
            c_ref, c2_ref = x, _ln(torch, x, g1.double(), b1.double())
        C, C2 = keng.debug_gemm_ln(A, Wt, bias, r if epi != 8 else None, 0.5, epi, (g1, b1), (g2, b2) if epi == 7 else None, inplace=inplace)
        torch.cuda.synchronize()
        assert not torch.isnan(C2).any()
        assert (C.double() - c_ref).abs().max().item() < 3e-2
        assert (C2.double() - c2_ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("B,T,H,dh,wf,wb", [(2, 250, 4, 36, -1, 0), (1, 106, 4, 36, -1, 0), (3, 13, 4, 64, -1, 0), (1, 300, 2, 32, -1, 0),
                                            (2, 780, 4, 64, -1, 0), (1, 1, 4, 36, -1, 0), (2, 257, 4, 36, -1, 0), (2, 120, 4, 36, 36, 0),
                                            (1, 300, 4, 36, 36, 8), (32, 250, 4, 36, -1, 0)])
def test_attention_kernels_vs_fp64(keng, torch_mod, B, T, H, dh, wf, wb):
    """softmax(Q K^T) V per head (multihead_attention.py:151-188; optional ChunkConformer band, chunk_conformer_blocks.py:158-176):
    tcgen05 kernel and fp32 CUDA-core kernel against torch fp64, single and multiple key blocks, ragged T."""
    torch = torch_mod
    torch.manual_seed(B * 1000 + T)
    qkv = torch.randn(B * T, 3 * H * dh, device="cuda")
    qkv[:, :H * dh] *= 0.5
    q, k, v = (qkv[:, i * H * dh:(i + 1) * H * dh].reshape(B, T, H, dh).double() for i in range(3))
    s = torch.einsum("bnhd,bmhd->bhnm", q, k)
    if wf >= 0:
        i = torch.arange(T, device="cuda")[:, None]
        j = torch.arange(T, device="cuda")[None, :]
        lo = torch.clamp(torch.minimum(torch.clamp(i - wf, min=0), torch.tensor(T - wb, device="cuda")), min=0)
        hi = torch.clamp(torch.maximum(torch.minimum(i + wb, torch.tensor(T, device="cuda")), torch.tensor(wb, device="cuda")), max=T - 1)
        s = s.masked_fill(~((j >= lo) & (j <= hi)), float("-inf"))
    ref = torch.einsum("bhnm,bmhd->bnhd", torch.softmax(s, -1), v).reshape(B * T, H * dh)
    o_tc = keng.debug_attention(qkv, B, T, H, dh, True, wf, wb)
    o_32 = keng.debug_attention(qkv, B, T, H, dh, False, wf, wb)
    torch.cuda.synchronize()
    assert not torch.isnan(o_tc).any()
    assert (o_tc.double() - ref).abs().max().item() < 1e-2
    assert (o_32.double() - ref).abs().max().item() < 1e-4
    # cp.async staging (the engine's schedule: the QKV GEMM stores tf32 numbers, so staging neither rounds nor blocks on loads)
    qr = ((qkv.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    o_as = keng.debug_attention(qr, B, T, H, dh, 2, wf, wb)
    o_rn = keng.debug_attention(qr, B, T, H, dh, True, wf, wb)
    torch.cuda.synchronize()
    assert torch.equal(o_as, o_rn)              # identical tiles in shared memory -> identical results


@pytest.mark.parametrize("M", [8000, 300, 1, 129, 8064])
def test_pair_direct_projection_vs_fp64(keng, torch_mod, M):
    """Attention out-projection + residual + LayerNorm on the cluster-pair kernel (K = 144 split 72 / 72 across the two CTAs,
    partial sums exchanged through DSMEM) against torch fp64."""
    torch = torch_mod
    torch.manual_seed(M)
    D = 144
    X = torch.randn(M, D, device="cuda")
    W = torch.randn(D, D, device="cuda") / D ** 0.5
    bias = torch.randn(D, device="cuda") * 0.3
    g1, b1, g2, b2 = (torch.randn(D, device="cuda") for _ in range(4))
    resid = torch.randn(M, D, device="cuda") * 20
    x = resid.double() + 1.0 * (X.double() @ W.double().T + bias.double())
    for epi in (6, 7):
        if epi == 6:
            c_ref, c2_ref = x, _ln(torch, x, g1.double(), b1.double())
        else:
            c_ref = _ln(torch, x, g1.double(), b1.double())
            c2_ref = _ln(torch, c_ref, g2.double(), b2.double())
        C, C2 = keng.debug_pair_direct(X, W, bias, resid, 1.0, epi, (g1, b1), (g2, b2) if epi == 7 else None)
        torch.cuda.synchronize()
        assert not torch.isnan(C2).any()
        assert (C.double() - c_ref).abs().max().item() < 3e-2
        assert (C2.double() - c2_ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("B,T,D,K", [(3, 250, 144, 32), (2, 13, 256, 5), (1, 37, 144, 32), (4, 100, 64, 7), (2, 250, 288, 32)])
def test_depthwise_conv_kernels_vs_fp64(keng, torch_mod, B, T, D, K):
    """The conv module's depthwise convolution (SeparableConv1D depthwise part, 'same' = TF SAME_UPPER) alone: the packed
    fp32x2 register kernel (K = 32, even D), the scalar register kernel (K = 5) and the generic kernel, exact fp32 and with the
    tf32-rounded store the tensor-core schedule uses."""
    torch = torch_mod
    torch.manual_seed(B * 1000 + T + D + K)
    x = torch.randn(B, T, D, device="cuda")
    w = torch.randn(K, D, device="cuda") / K ** 0.5
    total = K - 1
    pad_left = total // 2
    xp = torch.nn.functional.pad(x.double(), (0, 0, pad_left, total - pad_left))
    ref = sum(xp[:, j:j + T, :] * w[j].double() for j in range(K))
    y = keng.debug_dwconv(x, w, pad_left)
    yr = keng.debug_dwconv(x, w, pad_left, round_tf32=True)
    torch.cuda.synchronize()
    assert (y.double() - ref).abs().max().item() < 2e-5
    assert (yr.double() - ref).abs().max().item() < 2e-5 + ref.abs().max().item() * 2.0 ** -11
    assert (yr.view(torch.int32) & 0x1FFF).abs().max().item() == 0          # stored values are exact tf32 numbers
    causal = keng.debug_dwconv(x, w, K - 1)                                  # the ChunkConformer's causal variant (pad_left = K - 1)
    xc = torch.nn.functional.pad(x.double(), (0, 0, K - 1, 0))
    refc = sum(xc[:, j:j + T, :] * w[j].double() for j in range(K))
    assert (causal.double() - refc).abs().max().item() < 2e-5


@pytest.mark.parametrize("B,T", [(2, 1000), (3, 421), (1, 7), (2, 50)])
def test_subsampling_convs_vs_fp64(torch_mod, B, T):
    """conv1 (3x3 s2 'same', 1 -> D, ReLU; CUDA cores) + conv2 (3x3 s2 'same', D -> D, ReLU; implicit GEMM on tcgen05 with the
    4-D strided TMA A operand, a_mode 1) against torch conv2d in fp64, tf32 engine and exact-fp32 engine.  Odd T exercises
    every TF SAME_UPPER padding branch."""
    torch = torch_mod
    from tensorflowasr_b200 import engine as E, weights as W
    ge, re_, gc, rc = W.random_model(5, num_blocks=1)
    D = ge.dmodel
    torch.manual_seed(B * 31 + T)
    mel = (torch.randn(B, T, 80, device="cuda") * 20 - 40)
    w1 = torch.from_numpy(re_["sub.conv1.w"]).cuda().double()            # [3, 3, 1, D] HWIO
    b1 = torch.from_numpy(re_["sub.conv1.b"]).cuda().double()
    w2 = torch.from_numpy(re_["sub.conv2.w"]).cuda().double()            # [3, 3, D, D]
    b2 = torch.from_numpy(re_["sub.conv2.b"]).cuda().double()

    def same_conv(x, w, b):                                              # x [B, C, H, W]; TF 'same' stride 2: extra pad at the end
        H, Wd = x.shape[2], x.shape[3]
        ph = max((-(-H // 2) - 1) * 2 + 3 - H, 0)
        pw = max((-(-Wd // 2) - 1) * 2 + 3 - Wd, 0)
        x = torch.nn.functional.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        return torch.relu(torch.nn.functional.conv2d(x, w.permute(3, 2, 0, 1), b, stride=2))

    ref = same_conv(same_conv(mel.double()[:, None], w1, b1), w2, b2).permute(0, 2, 3, 1)     # [B, T2, F2, D]
    import os
    tf32_tol = 3e-2 * max(1.0, ref.abs().max().item() / 100)
    # exact fp32 kernels; tf32 two-kernel path with the conv1 map in fp16 + kind::f16 conv2 (default) and in tf32-rounded fp32 +
    # kind::tf32 (B200ASR_NO_CONV_F16=1): both 4-D strided TMA; tf32 fused conv1 -> conv2 kernel (opt-in switch)
    errs = {}
    for prec, fused, no_f16, tol in ((1, "0", "0", 1e-3), (0, "0", "0", tf32_tol), (0, "0", "1", tf32_tol), (0, "1", "0", tf32_tol)):
        os.environ["B200ASR_FUSED_SUB"] = fused
        os.environ["B200ASR_NO_CONV_F16"] = no_f16
        try:
            e = E.Engine(ge, re_, gc, rc, precision=prec, use_cuda_graph=False)
        finally:
            os.environ.pop("B200ASR_FUSED_SUB", None)
            os.environ.pop("B200ASR_NO_CONV_F16", None)
        got = e.debug_subsample_convs(mel)
        torch.cuda.synchronize()
        assert tuple(got.shape) == tuple(ref.shape)
        err = (got.double() - ref).abs().max().item()
        errs[(prec, fused, no_f16)] = err
        assert err < tol, (prec, fused, no_f16, err, ref.abs().max().item())
        e.close()
    # fp16 operands carry the same 11-bit significand as tf32-rounded ones: the two tensor-core variants err alike
    assert errs[(0, "0", "0")] < 1.5 * errs[(0, "0", "1")] + 1e-4, errs
