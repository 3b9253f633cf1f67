"""ChunkConformer (causal chunk streaming with state caches) on the B200 path: weight packing, ctypes binding of the stream-state
C ABI (include/b200asr.h) and the reference-facing runner surface.

Reference: asr/models/chunk_conformer_blocks.py (`ChunkConformer`: init_picker_caches :777-786, init_decoder_caches :788-792,
predict :798-805, picker_stream_predict :807-824, decoder_stream_predict :826-837, feature_pick :913-999) and its driver
test_chunk_asr.py:47-139 (`ASR.stream_call`).  The reference ships no ChunkConformer weights: `random_chunk_weights` builds
random-initialised weights of the architecture (asr/configs/chunk_conformerS.yml) for benchmarks; a trained model's variables map
onto the same names.

Tensor names (float32): fe.window [1024], fe.mel [513, 80]; sub.conv1.w [3,3,1,D] .b, sub.conv2.w [3,3,D,D] .b, sub.lin.w [F2*D, D] .b;
per block <stack>.<i>. (stack in enc / picker / helper / dec): ffn{1,2}.ln.{g,b}, ffn{1,2}.{w1 [D,4D], b1, w2 [4D,D], b2},
mhsa.ln.{g,b}, mhsa.{wq,wk,wv} [D,H,dh] + .{bq,bk,bv} [H,dh] (Keras MultiHeadAttention layout), mhsa.wo [H,dh,D], mhsa.bo,
conv.ln.{g,b}, conv.pw1.{w [D,2D], b}, conv.dw.w [K,D], conv.pw.{w [D,2D], b}, conv.bn.{scale,shift} [2D] (eval-mode BatchNorm
folded to an affine), conv.pw2.{w [2D,D], b}, ln.{g,b}; picker.proj / picker.fc / dec.proj / dec.fc .{w [in,out], b}.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import engine as E
from . import weights as W


@dataclass
class ChunkGeometry:
    """asr/configs/chunk_conformerS.yml model_config + chunk_data.yml speech_config."""
    dmodel: int = 144
    num_heads: int = 4
    head_size: int = 36
    kernel_size: int = 32
    ff_dim: int = 576
    enc_blocks: int = 15
    picker_blocks: int = 1
    helper_blocks: int = 2
    dec_blocks: int = 1
    win_front: int = 36
    picker_back: int = 0
    dec_back: int = 8
    phone_classes: int = 277
    txt_classes: int = 9171
    n_mels: int = 80
    n_dft: int = 1024
    hop: int = 160
    chunk_num: int = 16          # mel frames per step: 16 = 160 ms (shipped config), 32 = 320 ms (BASELINE.json configs[3])
    reduction: int = 4
    ln_eps: float = 1e-3

    @property
    def frames_per_step(self) -> int:
        return self.chunk_num // self.reduction

    @property
    def samples_per_step(self) -> int:
        return self.chunk_num * self.hop

    @property
    def f2(self) -> int:
        return (((self.n_mels + 4 - 3) // 2 + 1) - 3) // 2 + 1


class ChunkConfig(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32),
                ("dmodel", ctypes.c_int32), ("num_heads", ctypes.c_int32), ("head_size", ctypes.c_int32), ("kernel_size", ctypes.c_int32),
                ("ff_dim", ctypes.c_int32),
                ("enc_blocks", ctypes.c_int32), ("picker_blocks", ctypes.c_int32), ("helper_blocks", ctypes.c_int32), ("dec_blocks", ctypes.c_int32),
                ("win_front", ctypes.c_int32), ("picker_back", ctypes.c_int32), ("dec_back", ctypes.c_int32),
                ("phone_classes", ctypes.c_int32), ("txt_classes", ctypes.c_int32),
                ("n_mels", ctypes.c_int32), ("n_dft", ctypes.c_int32), ("hop", ctypes.c_int32), ("chunk_num", ctypes.c_int32),
                ("reduction", ctypes.c_int32),
                ("ln_eps", ctypes.c_float), ("use_cuda_graph", ctypes.c_int32), ("reserved", ctypes.c_int32 * 8)]


# --------------------------------------------------------------------------------------------------------- weights
def random_chunk_weights(seed: int, fe_raw: Dict[str, np.ndarray], geo: ChunkGeometry = ChunkGeometry(), scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Random-initialised ChunkConformer weights (float32) with the tensor names listed in the module docstring.  fe_raw supplies the
    Hann window and mel matrix ('fe.window', 'fe.mel': the ones baked into the reference's shipped ONNX graphs)."""
    rng = np.random.default_rng(seed)
    D, H, dh, K = geo.dmodel, geo.num_heads, geo.head_size, geo.kernel_size
    raw: Dict[str, np.ndarray] = {"fe.window": np.asarray(fe_raw["fe.window"], np.float32), "fe.mel": np.asarray(fe_raw["fe.mel"], np.float32)}

    def put(name, arr):
        raw[name] = np.asarray(arr, dtype=np.float32)

    def dense(name, i, o):
        put(name + ".w", rng.standard_normal((i, o)) * scale / np.sqrt(i))
        put(name + ".b", rng.standard_normal(o) * 0.1)

    def ln(name):
        put(name + ".g", 1.0 + 0.1 * rng.standard_normal(D))
        put(name + ".b", 0.1 * rng.standard_normal(D))

    def block(p):
        for f in ("ffn1", "ffn2"):
            ln(p + f + ".ln")
            put(p + f + ".w1", rng.standard_normal((D, geo.ff_dim)) * scale / np.sqrt(D))
            put(p + f + ".b1", rng.standard_normal(geo.ff_dim) * 0.1)
            put(p + f + ".w2", rng.standard_normal((geo.ff_dim, D)) * scale / np.sqrt(geo.ff_dim))
            put(p + f + ".b2", rng.standard_normal(D) * 0.1)
        ln(p + "mhsa.ln")
        for n in ("q", "k", "v"):
            put(p + f"mhsa.w{n}", rng.standard_normal((D, H, dh)) * scale / np.sqrt(D))
            put(p + f"mhsa.b{n}", rng.standard_normal((H, dh)) * 0.1)
        put(p + "mhsa.wo", rng.standard_normal((H, dh, D)) * scale / np.sqrt(H * dh))
        put(p + "mhsa.bo", rng.standard_normal(D) * 0.1)
        ln(p + "conv.ln")
        put(p + "conv.pw1.w", rng.standard_normal((D, 2 * D)) * scale / np.sqrt(D))
        put(p + "conv.pw1.b", rng.standard_normal(2 * D) * 0.1)
        put(p + "conv.dw.w", rng.standard_normal((K, D)) * scale / np.sqrt(K))
        put(p + "conv.pw.w", rng.standard_normal((D, 2 * D)) * scale / np.sqrt(D))
        put(p + "conv.pw.b", rng.standard_normal(2 * D) * 0.1)
        put(p + "conv.bn.scale", 1.0 + 0.1 * rng.standard_normal(2 * D))
        put(p + "conv.bn.shift", 0.1 * rng.standard_normal(2 * D))
        put(p + "conv.pw2.w", rng.standard_normal((2 * D, D)) * scale / np.sqrt(2 * D))
        put(p + "conv.pw2.b", rng.standard_normal(D) * 0.1)
        ln(p + "ln")

    put("sub.conv1.w", rng.standard_normal((3, 3, 1, D)) * scale / 3.0)
    put("sub.conv1.b", rng.standard_normal(D) * 0.1)
    put("sub.conv2.w", rng.standard_normal((3, 3, D, D)) * scale / np.sqrt(9 * D))
    put("sub.conv2.b", rng.standard_normal(D) * 0.1)
    dense("sub.lin", geo.f2 * D, D)
    for stack, n in (("enc", geo.enc_blocks),):
        for i in range(n):
            block(f"{stack}.{i}.")
    dense("picker.proj", D, D)
    for i in range(geo.picker_blocks):
        block(f"picker.{i}.")
    dense("picker.fc", D, geo.phone_classes)
    for i in range(geo.helper_blocks):
        block(f"helper.{i}.")
    dense("dec.proj", D, D)
    for i in range(geo.dec_blocks):
        block(f"dec.{i}.")
    dense("dec.fc", D, geo.txt_classes)
    return raw


def _pad_head(w: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Dense head [D, V] -> device layout [Vpad, D] with V padded to a multiple of 4 (tensor-map / float4 alignment): zero weight rows
    and a -1e30 bias, so a padded class can never win an argmax; the binding strips the padding from what it returns."""
    V = w.shape[1]
    Vp = (V + 3) // 4 * 4
    wt = np.zeros((Vp, w.shape[0]), np.float32)
    wt[:V] = w.T
    bb = np.full(Vp, -1e30, np.float32)
    bb[:V] = b
    return wt, bb


def chunk_device_tensors(raw: Dict[str, np.ndarray], geo: ChunkGeometry) -> Dict[str, np.ndarray]:
    """Device layout: every GEMM operand [N, K] K-major and rounded to nearest tf32 (weights.round_to_tf32), BatchNorm folded, GLU halves
    pair-interleaved, 1/sqrt(dh) folded into Wq AND bq, q|k|v biases concatenated (weights._pack_block does the shared part)."""
    D, H, dh = geo.dmodel, geo.num_heads, geo.head_size
    out: Dict[str, np.ndarray] = {}
    f32 = {k: np.asarray(v, np.float32) for k, v in raw.items()}
    out["fe.window"] = f32["fe.window"]
    out["fe.mel"] = f32["fe.mel"]
    out["sub.conv1.w"] = f32["sub.conv1.w"].reshape(9, D)
    out["sub.conv1.b"] = f32["sub.conv1.b"]
    out["sub.conv2.w"] = f32["sub.conv2.w"].transpose(3, 0, 1, 2).reshape(D, 9 * D)
    out["sub.conv2.b"] = f32["sub.conv2.b"]
    out["sub.lin.w"] = f32["sub.lin.w"].T
    out["sub.lin.b"] = f32["sub.lin.b"]
    stacks = (("enc", geo.enc_blocks), ("picker", geo.picker_blocks), ("helper", geo.helper_blocks), ("dec", geo.dec_blocks))
    for stack, n in stacks:
        for i in range(n):
            p = f"{stack}.{i}."
            blk = {k: v for k, v in f32.items() if k.startswith(p)}
            # Keras MultiHeadAttention kernels are [D, H, dh]; weights._pack_block expects the offline layout [H, D, dh]
            for nme in ("q", "k", "v"):
                blk[p + f"mhsa.w{nme}"] = f32[p + f"mhsa.w{nme}"].transpose(1, 0, 2)
            W._pack_block(blk, p, p, out)
            s = np.float32(1.0 / np.sqrt(np.float32(dh)))
            out[p + "mhsa.bqkv"] = np.concatenate([f32[p + "mhsa.bq"].reshape(H * dh) * s, f32[p + "mhsa.bk"].reshape(H * dh),
                                                    f32[p + "mhsa.bv"].reshape(H * dh)])
    out["picker.proj.w"] = f32["picker.proj.w"].T
    out["picker.proj.b"] = f32["picker.proj.b"]
    out["picker.fc.w"], out["picker.fc.b"] = _pad_head(f32["picker.fc.w"], f32["picker.fc.b"])
    out["dec.proj.w"] = f32["dec.proj.w"].T
    out["dec.proj.b"] = f32["dec.proj.b"]
    out["dec.fc.w"], out["dec.fc.b"] = _pad_head(f32["dec.fc.w"], f32["dec.fc.b"])
    out = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}
    extra = ("sub.conv2.w", "sub.lin.w", "picker.proj.w", "picker.fc.w", "dec.proj.w", "dec.fc.w")
    for k in out:
        if k.endswith(W._TC_OPERAND_SUFFIXES) or k in extra:
            out[k] = W.round_to_tf32(out[k])
    return out


def pack_chunk_blob(raw: Dict[str, np.ndarray], geo: ChunkGeometry) -> bytes:
    return W.pack_blob(chunk_device_tensors(raw, geo))


# --------------------------------------------------------------------------------------------------------- binding
def _bind(lib):
    if getattr(lib, "_chunk_bound", False):
        return lib
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.b200asr_chunk_create.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(ChunkConfig), ci, ctypes.POINTER(vp)]
    lib.b200asr_stream_state_create.argtypes = [vp, ci, ctypes.POINTER(vp)]
    lib.b200asr_stream_state_reset.argtypes = [vp, vp]
    lib.b200asr_stream_state_destroy.argtypes = [vp, vp]
    lib.b200asr_stream_step.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.b200asr_stream_feature_pick.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, vp, vp, ctypes.POINTER(ctypes.c_int32), vp]
    lib.b200asr_stream_decoder_step.argtypes = [vp, vp, vp, ci, vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), vp]
    lib.b200asr_stream_decoder_rows.argtypes = [vp, vp, ci]
    lib._chunk_bound = True
    return lib


class StreamState:
    """Caches of B lockstep streams (b200asr_stream): the device-side equivalent of the reference's caches lists."""

    def __init__(self, engine: "ChunkEngine", B: int):
        self.engine, self.B = engine, B
        h = ctypes.c_void_p()
        engine._check(engine.lib.b200asr_stream_state_create(engine._h, int(B), ctypes.byref(h)), "b200asr_stream_state_create")
        self._st = h

    def reset(self):
        self.engine._check(self.engine.lib.b200asr_stream_state_reset(self.engine._h, self._st), "b200asr_stream_state_reset")

    def close(self):
        if getattr(self, "_st", None) and getattr(self.engine, "_h", None):
            self.engine.lib.b200asr_stream_state_destroy(self.engine._h, self._st)
        self._st = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ChunkEngine:
    """One ChunkConformer handle on one GPU (b200asr_chunk_create).  No CPU fallback."""

    def __init__(self, raw: Dict[str, np.ndarray], geo: ChunkGeometry = ChunkGeometry(), device: int = 0, use_cuda_graph: bool = True):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("b200asr needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _bind(E.load_library())
        self.geo, self.device = geo, device
        blob = pack_chunk_blob(raw, geo)
        cfg = ChunkConfig()
        cfg.abi_version = self.lib.b200asr_abi_version()
        for f in ("dmodel", "num_heads", "head_size", "kernel_size", "ff_dim", "enc_blocks", "picker_blocks", "helper_blocks", "dec_blocks",
                  "win_front", "picker_back", "dec_back", "phone_classes", "txt_classes", "n_mels", "n_dft", "hop", "chunk_num", "reduction"):
            setattr(cfg, f, int(getattr(geo, f)))
        cfg.ln_eps = geo.ln_eps
        cfg.use_cuda_graph = int(bool(use_cuda_graph))
        h = ctypes.c_void_p()
        torch.cuda.set_device(device)
        buf = ctypes.create_string_buffer(blob, len(blob))
        rc = self.lib.b200asr_chunk_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.byref(cfg), device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("b200asr_chunk_create: " + self.lib.b200asr_last_error(None).decode(errors="replace"))
        self._h = h

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
        import torch
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        import torch
        return torch.device("cuda", self.device)

    @property
    def launch_count(self) -> int:
        return int(self.lib.b200asr_launch_count(self._h))

    def new_state(self, B: int) -> StreamState:
        return StreamState(self, B)

    # ------------------------------------------------------------------------------------------------ steps (device tensors)
    def picker_step(self, state: StreamState, wav_chunk, phone_logits=None, hidden=None):
        """wav_chunk [B, chunk_num*hop] (cuda float32) -> (phone logits [B, T, phone_classes], hidden [B, T, D]), T = chunk_num / 4."""
        import torch
        g = self.geo
        B, T = state.B, g.frames_per_step
        assert tuple(wav_chunk.shape) == (B, g.samples_per_step), (tuple(wav_chunk.shape), (B, g.samples_per_step))
        if phone_logits is None:
            phone_logits = torch.empty((B, T, g.phone_classes), device=self._dev(), dtype=torch.float32)
        if hidden is None:
            hidden = torch.empty((B, T, g.dmodel), device=self._dev(), dtype=torch.float32)
        self._check(self.lib.b200asr_stream_step(self._h, state._st, wav_chunk.data_ptr(), phone_logits.data_ptr(), hidden.data_ptr(),
                                                 self._stream()), "b200asr_stream_step")
        return phone_logits, hidden

    def feature_pick(self, hidden, phone_logits, blank: Optional[int] = None, want_logits: bool = True):
        """-> (feats [B, n, D], picked logits [B, n, V] or None, counts [B]) with n = the largest number of non-blank frames of any
        stream (synchronises: the shape is data dependent, exactly as in the reference)."""
        import torch
        B, T, D = hidden.shape
        V = phone_logits.shape[-1]
        blank = V - 1 if blank is None else int(blank)
        feats = torch.empty((B, T, D), device=self._dev(), dtype=torch.float32)
        picked = torch.empty((B, T, V), device=self._dev(), dtype=torch.float32) if want_logits else None
        counts = torch.empty((B,), device=self._dev(), dtype=torch.int32)
        n = ctypes.c_int32(0)
        self._check(self.lib.b200asr_stream_feature_pick(self._h, hidden.contiguous().data_ptr(), phone_logits.contiguous().data_ptr(), B, T, V, blank,
                                                         feats.data_ptr(), picked.data_ptr() if picked is not None else None, counts.data_ptr(),
                                                         ctypes.byref(n), self._stream()), "b200asr_stream_feature_pick")
        n = int(n.value)
        return feats[:, :n].contiguous(), (picked[:, :n].contiguous() if picked is not None else None), counts

    def decoder_step(self, state: StreamState, feats):
        """feats [B, n, D] (n >= 1) -> (valid text logits [B, n_valid, txt_classes], unvalid text logits [B, n_rows - n_valid, txt_classes])."""
        import torch
        g = self.geo
        B, n, _ = feats.shape
        rows = int(self.lib.b200asr_stream_decoder_rows(self._h, state._st, int(n)))
        out = torch.empty((B, rows, g.txt_classes), device=self._dev(), dtype=torch.float32)
        n_rows, n_valid = ctypes.c_int32(0), ctypes.c_int32(0)
        self._check(self.lib.b200asr_stream_decoder_step(self._h, state._st, feats.contiguous().data_ptr(), int(n), out.data_ptr(), ctypes.byref(n_rows),
                                                         ctypes.byref(n_valid), self._stream()), "b200asr_stream_decoder_step")
        assert n_rows.value == rows
        return out[:, :n_valid.value], out[:, n_valid.value:]


# --------------------------------------------------------------------------------------------------------- reference-facing surface
class ChunkConformer:
    """The `runner` object of test_chunk_asr.py: same method names, argument order and return values as the reference's
    ChunkConformer (chunk_conformer_blocks.py:777-852), with torch CUDA tensors where the reference returns TF tensors.

    `caches` objects are opaque here (a StreamState per call site): the reference's caches are lists of TF tensors that the caller
    only ever passes back in, which is what these do.  A picker state and a decoder state may be the same object (one StreamState
    holds both groups of caches); init_decoder_caches therefore returns the state it is given or a fresh one."""

    def __init__(self, engine: ChunkEngine):
        self.engine = engine
        self.geo = engine.geo
        self.num_classes = engine.geo.phone_classes      # self.num_classes of the reference = phone classes (blank = num_classes - 1)

    def init_picker_caches(self, B: int) -> StreamState:
        return self.engine.new_state(B)

    def init_decoder_caches(self, B: int, state: Optional[StreamState] = None) -> StreamState:
        return state if state is not None else self.engine.new_state(B)

    def _as_chunk(self, wav_chunk):
        import torch
        if isinstance(wav_chunk, np.ndarray):
            wav_chunk = torch.from_numpy(np.ascontiguousarray(wav_chunk, dtype=np.float32))
        wav_chunk = wav_chunk.to(device=self.engine._dev(), dtype=torch.float32)
        if wav_chunk.dim() == 3:
            wav_chunk = wav_chunk[..., 0]
        S = self.geo.samples_per_step
        if wav_chunk.shape[1] < S:      # the last, ragged chunk of an utterance: zero-filled to a whole step
            wav_chunk = torch.nn.functional.pad(wav_chunk, (0, S - wav_chunk.shape[1]))
        return wav_chunk.contiguous()

    def picker_stream_predict(self, wav_chunk, caches: StreamState):
        """-> (valid phone logits [B, T, V], unvalid phone logits (zeros: the picker has no look-ahead, :655), valid hidden [B, T, D], caches)."""
        import torch
        logits, hidden = self.engine.picker_step(caches, self._as_chunk(wav_chunk))
        return logits, torch.zeros_like(logits), hidden, caches

    def feature_pick(self, hidden, phone_logits):
        feats, picked, _ = self.engine.feature_pick(hidden, phone_logits, blank=self.num_classes - 1)
        return feats, picked

    def decoder_stream_predict(self, features, caches: StreamState):
        """-> (valid text logits, unvalid text logits, caches)."""
        valid, unvalid = self.engine.decoder_step(caches, features)
        return valid, unvalid, caches

    def predict(self, wav):
        """Offline text logits [B, U, txt_classes] (:798-805), evaluated by streaming the utterance through fresh caches: the reference's
        own invariant (test_chunk_asr.py:57,123,139) is that the concatenated valid streaming outputs followed by the final unvalid ones
        equal the offline outputs."""
        import torch
        if isinstance(wav, np.ndarray):
            wav = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))
        if wav.dim() == 3:
            wav = wav[..., 0]
        B = wav.shape[0]
        st = self.engine.new_state(B)
        S = self.geo.samples_per_step
        outs: List = []
        last_unvalid = None
        for s in range(0, wav.shape[1], S):
            ph, _, hid, _ = self.picker_stream_predict(wav[:, s:s + S], st)
            feats, _ = self.feature_pick(hid, ph)
            if feats.shape[1] == 0:
                continue
            v, u, _ = self.decoder_stream_predict(feats, st)
            outs.append(v)
            last_unvalid = u
        st.close()
        if last_unvalid is not None:
            outs.append(last_unvalid)
        if not outs:
            return torch.zeros((B, 0, self.geo.txt_classes), device=self.engine._dev())
        return torch.cat(outs, dim=1)

    def stream_call(self, wav, on_partial=None):
        """The driver loop of test_chunk_asr.py:47-93: feed `wav` [1, L] chunk by chunk, greedy-decode the text logits seen so far after
        every chunk (valid frames + the current unvalid tail) and hand the running token ids to `on_partial`.  Returns the final ids."""
        import torch
        if isinstance(wav, np.ndarray):
            wav = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))
        if wav.dim() == 3:
            wav = wav[..., 0]
        if wav.dim() == 1:
            wav = wav[None]
        st = self.engine.new_state(wav.shape[0])
        S = self.geo.samples_per_step
        blank = self.geo.txt_classes - 1
        valid_ids: List[int] = []
        ids: List[int] = []
        for s in range(0, wav.shape[1], S):
            ph, _, hid, _ = self.picker_stream_predict(wav[:, s:s + S], st)
            feats, _ = self.feature_pick(hid, ph)
            if feats.shape[1] == 0:
                continue
            v, u, _ = self.decoder_stream_predict(feats, st)
            valid_ids += v[0].argmax(-1).tolist()
            frames = valid_ids + u[0].argmax(-1).tolist()
            ids, prev = [], None
            for t in frames:
                if t != prev and t != blank:
                    ids.append(int(t))
                prev = t
            if on_partial is not None:
                on_partial(ids)
        st.close()
        return ids
