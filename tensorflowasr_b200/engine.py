"""ctypes binding of libb200asr.so (include/b200asr.h).  PyTorch only supplies device memory and streams.

The product path has no CPU fallback: if the CUDA library is missing or no sm_100 GPU is present, constructing
an Engine raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import numpy as np

from . import weights as W

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200asr.so")

PRECISION_TF32 = 0
PRECISION_FP32 = 1


class Config(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32),
                ("dmodel", ctypes.c_int32), ("num_blocks", ctypes.c_int32), ("num_heads", ctypes.c_int32),
                ("head_size", ctypes.c_int32), ("kernel_size", ctypes.c_int32), ("ff_dim", ctypes.c_int32),
                ("ctc_blocks", ctypes.c_int32), ("ctc_kernel_size", ctypes.c_int32), ("vocab", ctypes.c_int32),
                ("n_mels", ctypes.c_int32), ("n_dft", ctypes.c_int32), ("hop", ctypes.c_int32),
                ("ln_eps", ctypes.c_float),
                ("chunk_samples", ctypes.c_int32), ("precision", ctypes.c_int32), ("use_cuda_graph", ctypes.c_int32),
                ("tr_blocks", ctypes.c_int32), ("tr_kernel_size", ctypes.c_int32), ("tr_inp_classes", ctypes.c_int32),
                ("tr_vocab", ctypes.c_int32),
                ("reserved", ctypes.c_int32 * 4)]


_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.b200asr_abi_version.restype = ci
    lib.b200asr_create.restype = ci
    lib.b200asr_create.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(Config), ci, ctypes.POINTER(vp)]
    lib.b200asr_destroy.argtypes = [vp]
    lib.b200asr_last_error.restype = ctypes.c_char_p
    lib.b200asr_last_error.argtypes = [vp]
    lib.b200asr_out_frames.restype = ci
    lib.b200asr_out_frames.argtypes = [vp, ci]
    lib.b200asr_reserve.argtypes = [vp, ci, ci]
    lib.b200asr_encode.argtypes = [vp, vp, ci, ci, vp, vp]
    lib.b200asr_mel.argtypes = [vp, vp, ci, ci, vp, vp]
    lib.b200asr_ctc_logits.argtypes = [vp, vp, ci, ci, vp, vp]
    lib.b200asr_ctc_greedy.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, vp, vp]
    lib.b200asr_ctc_beam.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, cf, vp, vp, vp, vp]
    lib.b200asr_ctc_beam_probs.argtypes = lib.b200asr_ctc_beam.argtypes
    lib.b200asr_translate.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp]
    lib.b200asr_recognize.argtypes = [vp, vp, ci, ci, vp, vp, vp]
    lib.b200asr_recognize_lengths.argtypes = [vp, vp, vp, ci, ci, vp, vp, vp]
    lib.b200asr_recognize_host.argtypes = [vp, vp, ci, ci, vp, vp, vp]
    lib.b200asr_recognize_host_submit.argtypes = [vp, ci, vp, ci, ci, vp, vp]
    lib.b200asr_recognize_host_collect.argtypes = [vp, ci]
    lib.b200asr_time_stage.argtypes = [vp, ci, ci, ci, ci, vp, ctypes.POINTER(cf), ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_double)]
    lib.b200asr_debug_gemm.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, cf, ci, ci, vp]
    lib.b200asr_debug_gemm_ln.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, cf, ci, vp, vp, vp, vp, cf, vp]
    lib.b200asr_debug_attention.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
    lib.b200asr_debug_chain.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, cf, ci, vp, vp, vp, vp, cf, vp]
    lib.b200asr_debug_chain_pair.argtypes = lib.b200asr_debug_chain.argtypes
    lib.b200asr_debug_dwconv.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
    lib.b200asr_debug_subsample_convs.argtypes = [vp, vp, ci, ci, vp, vp]
    lib.b200asr_debug_encode_taps.argtypes = [vp, vp, ci, ci, vp, ci, vp]
    lib.b200asr_vad_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, cf, ci, ctypes.POINTER(vp)]
    lib.b200asr_vad_infer.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    lib.b200asr_punc_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, cf, ci, ctypes.POINTER(vp)]
    lib.b200asr_punc_infer.argtypes = [vp, vp, ci, vp, vp]
    lib.b200asr_launch_count.restype = ctypes.c_int64
    lib.b200asr_launch_count.argtypes = [vp]
    _lib = lib
    return lib


def _torch():
    import torch
    return torch


class Engine:
    """One b200asr handle on one GPU."""

    def __init__(self, enc_geo: W.ModelGeometry, enc_raw: Dict[str, np.ndarray],
                 ctc_geo: Optional[W.ModelGeometry] = None, ctc_raw: Optional[Dict[str, np.ndarray]] = None,
                 device: int = 0, precision: int = PRECISION_TF32, chunk_samples: int = 0, use_cuda_graph: bool = True,
                 tr_geo: Optional[W.ModelGeometry] = None, tr_raw: Optional[Dict[str, np.ndarray]] = None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("b200asr needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.device = device
        self.enc_geo, self.ctc_geo = enc_geo, ctc_geo
        if ctc_geo is not None:
            # the C config carries one block geometry: the CTC decoder's blocks run with the encoder's head split and widths
            for f in ("dmodel", "num_heads", "head_size", "ff_dim"):
                if getattr(ctc_geo, f) != getattr(enc_geo, f):
                    raise ValueError(f"CTC decoder {f}={getattr(ctc_geo, f)} differs from the encoder's {getattr(enc_geo, f)}: "
                                     "b200asr_config has a single block geometry")
        self.tr_geo = tr_geo
        if tr_geo is not None:
            for f in ("dmodel", "num_heads", "head_size", "ff_dim"):
                if getattr(tr_geo, f) != getattr(enc_geo, f):
                    raise ValueError(f"translator {f}={getattr(tr_geo, f)} differs from the encoder's {getattr(enc_geo, f)}")
        blob = W.pack_blob(W.device_tensors(enc_geo, enc_raw, ctc_geo, ctc_raw, round_tf32=(int(precision) == PRECISION_TF32),
                                            tr_geo=tr_geo, tr_raw=tr_raw))
        cfg = Config()
        cfg.abi_version = self.lib.b200asr_abi_version()
        cfg.dmodel, cfg.num_blocks, cfg.num_heads = enc_geo.dmodel, enc_geo.num_blocks, enc_geo.num_heads
        cfg.head_size, cfg.kernel_size, cfg.ff_dim = enc_geo.head_size, enc_geo.kernel_size, enc_geo.ff_dim
        cfg.ctc_blocks = ctc_geo.num_blocks if ctc_geo else 0
        cfg.ctc_kernel_size = ctc_geo.kernel_size if ctc_geo else 0
        cfg.vocab = ctc_geo.vocab if ctc_geo else 0
        cfg.n_mels, cfg.n_dft, cfg.hop, cfg.ln_eps = enc_geo.n_mels, enc_geo.n_dft, enc_geo.hop, enc_geo.ln_eps
        cfg.chunk_samples, cfg.precision, cfg.use_cuda_graph = int(chunk_samples), int(precision), int(bool(use_cuda_graph))
        if tr_geo is not None:
            cfg.tr_blocks, cfg.tr_kernel_size = tr_geo.num_blocks, tr_geo.kernel_size
            cfg.tr_inp_classes, cfg.tr_vocab = int(tr_raw["tr.emb"].shape[0]), tr_geo.vocab
        self.cfg = cfg
        h = ctypes.c_void_p()
        torch.cuda.set_device(device)
        buf = ctypes.create_string_buffer(blob, len(blob))
        rc = self.lib.b200asr_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(cfg), device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("b200asr_create: " + self.lib.b200asr_last_error(None).decode(errors="replace"))
        self._h = h

    # ------------------------------------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200asr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what}: " + self.lib.b200asr_last_error(self._h).decode(errors="replace"))

    def _stream(self) -> int:
        return int(_torch().cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        return _torch().device("cuda", self.device)

    def out_frames(self, num_samples: int) -> int:
        return int(self.lib.b200asr_out_frames(self._h, int(num_samples)))

    def reserve(self, B: int, L: int):
        self._check(self.lib.b200asr_reserve(self._h, int(B), int(L)), "b200asr_reserve")

    @property
    def launch_count(self) -> int:
        return int(self.lib.b200asr_launch_count(self._h))

    def _as_wav(self, wav):
        torch = _torch()
        if isinstance(wav, np.ndarray):
            wav = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))
        wav = wav.to(device=self._dev(), dtype=torch.float32)
        if wav.dim() == 3:
            wav = wav[..., 0]
        if wav.dim() == 1:
            wav = wav[None, :]
        return wav.contiguous()

    # ------------------------------------------------------------------------------------------------ ops
    def mel(self, wav):
        torch = _torch()
        wav = self._as_wav(wav)
        B, L = wav.shape
        T = -(-L // self.enc_geo.hop)
        out = torch.empty((B, T, self.enc_geo.n_mels), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_mel(self._h, wav.data_ptr(), B, L, out.data_ptr(), self._stream()), "b200asr_mel")
        return out

    def encode(self, wav, out=None):
        """wav [B, L] (or [B, L, 1]) -> encoder states [B, T', D] on the GPU."""
        torch = _torch()
        wav = self._as_wav(wav)
        B, L = wav.shape
        Tp = self.out_frames(L)
        if out is None:
            out = torch.empty((B, Tp, self.enc_geo.dmodel), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_encode(self._h, wav.data_ptr(), B, L, out.data_ptr(), self._stream()), "b200asr_encode")
        return out

    def ctc_logits(self, enc, out=None):
        torch = _torch()
        if isinstance(enc, np.ndarray):
            enc = torch.from_numpy(np.ascontiguousarray(enc, dtype=np.float32))
        enc = enc.to(device=self._dev(), dtype=torch.float32).contiguous()
        B, Tp, _ = enc.shape
        if out is None:
            out = torch.empty((B, Tp, self.ctc_geo.vocab), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_ctc_logits(self._h, enc.data_ptr(), B, Tp, out.data_ptr(), self._stream()),
                    "b200asr_ctc_logits")
        return out

    def ctc_greedy(self, logits, lengths=None, blank: Optional[int] = None):
        """logits [B, T, V] -> (ids [B, T] int32 padded with -1, lengths [B] int32)."""
        torch = _torch()
        if isinstance(logits, np.ndarray):
            logits = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32))
        logits = logits.to(device=self._dev(), dtype=torch.float32).contiguous()
        B, T, V = logits.shape
        blank = V - 1 if blank is None else int(blank)
        ids = torch.empty((B, T), device=self._dev(), dtype=torch.int32)
        lens = torch.empty((B,), device=self._dev(), dtype=torch.int32)
        lptr = None
        if lengths is not None:
            lengths = torch.as_tensor(lengths, dtype=torch.int32).to(self._dev()).contiguous()
            lptr = lengths.data_ptr()
        self._check(self.lib.b200asr_ctc_greedy(self._h, logits.data_ptr(), lptr, B, T, V, blank, ids.data_ptr(),
                                                lens.data_ptr(), self._stream()), "b200asr_ctc_greedy")
        return ids, lens

    def ctc_beam(self, logits, beam: int, lengths=None, blank: Optional[int] = None, cutoff_top_n: int = 40,
                 cutoff_prob: float = 1.0, probs: bool = False):
        """Prefix beam search (no scorer).  -> (ids [B, beam, T] int32 -1 padded, lens [B, beam], scores [B, beam]).
        probs=True: `logits` holds probabilities (the reference decoder's own input); scores then equal the reference's to the last bit or two."""
        torch = _torch()
        if isinstance(logits, np.ndarray):
            logits = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32))
        logits = logits.to(device=self._dev(), dtype=torch.float32).contiguous()
        B, T, V = logits.shape
        blank = V - 1 if blank is None else int(blank)
        ids = torch.empty((B, beam, T), device=self._dev(), dtype=torch.int32)
        lens = torch.empty((B, beam), device=self._dev(), dtype=torch.int32)
        scores = torch.empty((B, beam), device=self._dev(), dtype=torch.float32)
        lptr = None
        if lengths is not None:
            lengths = torch.as_tensor(lengths, dtype=torch.int32).to(self._dev()).contiguous()
            lptr = lengths.data_ptr()
        fn = self.lib.b200asr_ctc_beam_probs if probs else self.lib.b200asr_ctc_beam
        self._check(fn(self._h, logits.data_ptr(), lptr, B, T, V, blank, int(beam), int(cutoff_top_n), float(cutoff_prob), ids.data_ptr(),
                       lens.data_ptr(), scores.data_ptr(), self._stream()), "b200asr_ctc_beam")
        return ids, lens, scores

    def translate(self, ids, enc):
        """Translator: ids [B, U] int32 (greedy phone ids, zero padded), enc [B, T', D] -> character logits [B, U, tr_vocab]."""
        torch = _torch()
        if self.tr_geo is None:
            raise RuntimeError("this engine was built without translator weights")
        ids = torch.as_tensor(np.asarray(ids) if not hasattr(ids, "device") else ids).to(device=self._dev(), dtype=torch.int32).contiguous()
        if isinstance(enc, np.ndarray):
            enc = torch.from_numpy(np.ascontiguousarray(enc, dtype=np.float32))
        enc = enc.to(device=self._dev(), dtype=torch.float32).contiguous()
        B, U = ids.shape
        Tp = enc.shape[1]
        out = torch.empty((B, U, self.tr_geo.vocab), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_translate(self._h, ids.data_ptr(), enc.data_ptr(), B, U, Tp, out.data_ptr(), self._stream()),
                    "b200asr_translate")
        return out

    def recognize(self, wav, ids=None, lens=None, frame_lengths=None):
        """wav [B, L] on the GPU -> greedy ids [B, T'] (-1 padded) + lengths, all on the GPU (no sync).  frame_lengths [B]
        (encoder frames, int32) = the `input_length` the reference hands to ctc_decode (am_tester.py:39); None = all frames."""
        torch = _torch()
        wav = self._as_wav(wav)
        B, L = wav.shape
        Tp = self.out_frames(L)
        if ids is None:
            ids = torch.empty((B, Tp), device=self._dev(), dtype=torch.int32)
        if lens is None:
            lens = torch.empty((B,), device=self._dev(), dtype=torch.int32)
        fl = None
        if frame_lengths is not None:
            frame_lengths = torch.as_tensor(frame_lengths, dtype=torch.int32).to(self._dev()).contiguous()
            fl = frame_lengths.data_ptr()
        self._check(self.lib.b200asr_recognize_lengths(self._h, wav.data_ptr(), fl, B, L, ids.data_ptr(), lens.data_ptr(),
                                                       self._stream()), "b200asr_recognize")
        return ids, lens

    def encode_taps(self, wav):
        """Test hook: residual stream after the subsampler and after every encoder block -> [1 + num_blocks, B, T', D]."""
        torch = _torch()
        wav = self._as_wav(wav)
        B, L = wav.shape
        Tp = self.out_frames(L)
        n = 1 + self.enc_geo.num_blocks
        taps = torch.empty((n, B, Tp, self.enc_geo.dmodel), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_debug_encode_taps(self._h, wav.data_ptr(), B, L, taps.data_ptr(), n, self._stream()),
                    "b200asr_debug_encode_taps")
        return taps

    def debug_dwconv(self, x, w, pad_left, round_tf32=False):
        """Test hook: depthwise conv x [B, T, D] (cuda), taps w [K, D] -> y [B, T, D]."""
        torch = _torch()
        B, T, D = x.shape
        y = torch.empty_like(x)
        self._check(self.lib.b200asr_debug_dwconv(self._h, x.data_ptr(), w.data_ptr(), y.data_ptr(), B, T, D, w.shape[0], int(pad_left),
                                                  int(bool(round_tf32)), self._stream()), "b200asr_debug_dwconv")
        return y

    def debug_subsample_convs(self, mel):
        """Test hook: mel [B, T, n_mels] (cuda) -> relu(conv2(relu(conv1(mel)))) [B, T2, F2, D]."""
        torch = _torch()
        B, T, F = mel.shape
        T1 = -(-T // 2); T2 = -(-T1 // 2)
        F1 = -(-F // 2); F2 = -(-F1 // 2)
        out = torch.empty((B, T2, F2, self.enc_geo.dmodel), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_debug_subsample_convs(self._h, mel.contiguous().data_ptr(), B, T, out.data_ptr(), self._stream()),
                    "b200asr_debug_subsample_convs")
        return out

    def debug_gemm(self, A, Wt, bias=None, resid=None, alpha=1.0, epilogue=0, tensor_cores=True):
        """Test hook: epilogue(A[M,K] @ Wt[N,K].T) on the GPU through one of the two GEMM kernels."""
        torch = _torch()
        M, K = A.shape
        N = Wt.shape[0]
        ldc = N // 2 if epilogue == 3 else N
        C = torch.empty((M, ldc), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_debug_gemm(self._h, A.data_ptr(), Wt.data_ptr(), bias.data_ptr() if bias is not None else None,
                                                resid.data_ptr() if resid is not None else None, C.data_ptr(), M, N, K, K, ldc,
                                                float(alpha), int(epilogue), int(bool(tensor_cores)), self._stream()),
                    "b200asr_debug_gemm")
        return C

    def debug_gemm_ln(self, A, Wt, bias, resid, alpha, epilogue, ln1, ln2=None, eps=1e-3, inplace=False):
        """Test hook: fused-LayerNorm epilogues of the tcgen05 kernel.  Returns (C, C2)."""
        torch = _torch()
        M, K = A.shape
        N = Wt.shape[0]
        C = resid if (inplace and resid is not None) else torch.empty((M, N), device=self._dev(), dtype=torch.float32)
        C2 = torch.zeros((M, N), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_debug_gemm_ln(
            self._h, A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), resid.data_ptr() if resid is not None else None, C.data_ptr(),
            C2.data_ptr(), M, N, K, float(alpha), int(epilogue), ln1[0].data_ptr(), ln1[1].data_ptr(),
            ln2[0].data_ptr() if ln2 else None, ln2[1].data_ptr() if ln2 else None, float(eps), self._stream()), "b200asr_debug_gemm_ln")
        return C, C2

    def debug_attention(self, qkv, B, T, H, dh, tensor_cores=True, win_front=-1, win_back=0):
        """Test hook: qkv [B*T, 3*H*dh] (torch cuda) -> attention output [B*T, H*dh]."""
        torch = _torch()
        out = torch.empty((B * T, H * dh), device=self._dev(), dtype=torch.float32)
        mode = 2 if tensor_cores == 2 else int(bool(tensor_cores))     # 2: tcgen05 kernel with cp.async staging (qkv already holds tf32 numbers)
        self._check(self.lib.b200asr_debug_attention(self._h, qkv.data_ptr(), out.data_ptr(), B, T, H, dh, int(win_front),
                                                     int(win_back), mode, self._stream()), "b200asr_debug_attention")
        return out

    def debug_pair_direct(self, X, W, bias, resid, alpha, epilogue, ln1, ln2=None, eps=1e-3):
        """Test hook: 144 -> 144 projection + residual + LayerNorm through the cluster-pair kernel (K split across the pair)."""
        torch = _torch()
        M, K = X.shape
        N = W.shape[0]
        C = torch.empty((M, N), device=self._dev(), dtype=torch.float32)
        C2 = torch.zeros((M, N), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_debug_chain_pair(
            self._h, X.data_ptr(), None, None, W.data_ptr(), bias.data_ptr(), resid.data_ptr(), C.data_ptr(), C2.data_ptr(), M, K, 0, N,
            float(alpha), int(epilogue), ln1[0].data_ptr(), ln1[1].data_ptr(), ln2[0].data_ptr() if ln2 else None,
            ln2[1].data_ptr() if ln2 else None, float(eps), self._stream()), "b200asr_debug_chain_pair")
        return C, C2

    def debug_chain(self, X, W1, b1, W2, b2, resid, alpha, epilogue, ln1, ln2=None, eps=1e-3, inplace=True, pair=False):
        """Test hook: chained FFN-style kernel (pair=True: the 2-CTA-cluster variant).  Returns (C, C2)."""
        torch = _torch()
        M, K1 = X.shape
        N1, N2 = W1.shape[0], W2.shape[0]
        C = resid if inplace else torch.empty((M, N2), device=self._dev(), dtype=torch.float32)
        C2 = torch.zeros((M, N2), device=self._dev(), dtype=torch.float32)
        fn = self.lib.b200asr_debug_chain_pair if pair else self.lib.b200asr_debug_chain
        self._check(fn(
            self._h, X.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), resid.data_ptr(), C.data_ptr(),
            C2.data_ptr(), M, K1, N1, N2, float(alpha), int(epilogue), ln1[0].data_ptr(), ln1[1].data_ptr(),
            ln2[0].data_ptr() if ln2 else None, ln2[1].data_ptr() if ln2 else None, float(eps), self._stream()), "b200asr_debug_chain")
        return C, C2

    STAGES = {"conv2": 0, "ffn_w1": 1, "ffn_w2": 2, "stft": 3, "sub_linear": 4, "attention": 5, "ctc_fc": 6, "ffn_chain": 7,
              "conv1": 8, "dwconv": 9, "qkv": 10}

    def time_stage(self, stage: str, B: int, L: int, iters: int = 20) -> Tuple[float, float, float]:
        """(ms per launch, algorithmic FLOPs per launch, algorithmic HBM bytes per launch) of one kernel timed alone
        with CUDA events on the current stream."""
        ms, fl, by = ctypes.c_float(0), ctypes.c_double(0), ctypes.c_double(0)
        self._check(self.lib.b200asr_time_stage(self._h, self.STAGES[stage], int(B), int(L), int(iters), self._stream(),
                                                ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by)), "b200asr_time_stage")
        return float(ms.value), float(fl.value), float(by.value)

    def recognize_host_submit(self, slot, wav_host, ids_host, lens_host):
        """Two-deep pipeline (b200asr_recognize_host_submit): pinned torch CPU tensors in / out, returns at once."""
        B, L = wav_host.shape
        self._check(self.lib.b200asr_recognize_host_submit(self._h, int(slot), wav_host.data_ptr(), B, L, ids_host.data_ptr(),
                                                           lens_host.data_ptr()), "b200asr_recognize_host_submit")

    def recognize_host_collect(self, slot):
        """Blocks until the slot's ids / lengths are in the host buffers given to recognize_host_submit."""
        self._check(self.lib.b200asr_recognize_host_collect(self._h, int(slot)), "b200asr_recognize_host_collect")

    def recognize_host(self, wav_host, ids_host=None, lens_host=None):
        """Host buffers in, host buffers out (torch CPU tensors, ideally pinned).  Synchronises."""
        torch = _torch()
        if isinstance(wav_host, np.ndarray):
            wav_host = torch.from_numpy(np.ascontiguousarray(wav_host, dtype=np.float32))
        if wav_host.dim() == 3:
            wav_host = wav_host[..., 0]
        wav_host = wav_host.contiguous()
        B, L = wav_host.shape
        Tp = self.out_frames(L)
        if ids_host is None:
            ids_host = torch.empty((B, Tp), dtype=torch.int32).pin_memory()
        if lens_host is None:
            lens_host = torch.empty((B,), dtype=torch.int32).pin_memory()
        self._check(self.lib.b200asr_recognize_host(self._h, wav_host.data_ptr(), B, L, ids_host.data_ptr(),
                                                    lens_host.data_ptr(), self._stream()), "b200asr_recognize_host")
        return ids_host, lens_host


def engine_from_onnx(model_dir: str, device: int = 0, precision: int = PRECISION_TF32, chunk_samples: int = 0,
                     use_cuda_graph: bool = True) -> Engine:
    """Build an Engine from the reference's deployment directory (encoder.onnx + ctc_model.onnx + translator.onnx when present), the
    same files Inference/PythonInference/asr/src/asr.py:22-25 loads."""
    ge, re_ = W.import_encoder(os.path.join(model_dir, "encoder.onnx"))
    gc, rc = W.import_ctc_model(os.path.join(model_dir, "ctc_model.onnx"))
    gt = rt = None
    if os.path.isfile(os.path.join(model_dir, "translator.onnx")):
        gt, rt = W.import_translator(os.path.join(model_dir, "translator.onnx"))
    return Engine(ge, re_, gc, rc, device=device, precision=precision, chunk_samples=chunk_samples,
                  use_cuda_graph=use_cuda_graph, tr_geo=gt, tr_raw=rt)
