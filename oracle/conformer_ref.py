"""TEST INFRASTRUCTURE ONLY (oracle) -- NumPy restatement of the reference's Conformer-CTC forward.

Every function cites the reference source it restates (paths relative to /root/reference).  The arithmetic
is done in `dtype` (float64 by default: an "exact" answer both the reference's fp32 graph and the B200 fp32 /
tf32 kernels are compared with).  Pinned against the reference itself: tests/test_oracle.py checks this file
against the shipped ONNX graphs run through the reference's vendored onnxruntime (oracle/ort_ref.py) on
asr/BAC009S0764W0121.wav and on seeded noise, stage by stage (mel, subsampling, block 0, encoder out, logits).

Parity status: PINNED (golden ids 669 82 103 78 247 56 71 573 386 82 30 213 496 reproduced; see tests/golden).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

Raw = Dict[str, np.ndarray]


# ----------------------------------------------------------------------------------------------------- helpers
def tf_same_pad(n_in: int, k: int, stride: int) -> Tuple[int, int, int]:
    """TensorFlow 'SAME' (= ONNX SAME_UPPER): returns (n_out, pad_before, pad_after)."""
    n_out = -(-n_in // stride)
    total = max((n_out - 1) * stride + k - n_in, 0)
    return n_out, total // 2, total - total // 2


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def swish(x):
    """tf.keras.activations.swish = x * sigmoid(x) (conformer_blocks.py:118-119)."""
    return x * sigmoid(x)


def layer_norm(x, g, b, eps=1e-3):
    """tf.keras.layers.LayerNormalization() defaults: last axis, epsilon=1e-3 (conformer_blocks.py:116,158,190,256)."""
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


# ----------------------------------------------------------------------------------------------------- frontend
def power_spectrogram(wav: np.ndarray, window: np.ndarray, n_dft=1024, hop=160, dtype=np.float64) -> np.ndarray:
    """Spectrogram._spectrogram_mono (asr/models/layers/time_frequency.py:100-122): two strided convs with
    cos/-sin x Hann kernels (backend.py:27-69), padding 'same', then re^2 + im^2.  wav [B, L] -> [B, T, 513].
    Evaluated with an FFT of the windowed frames, which is the same sum."""
    wav = np.asarray(wav, dtype=dtype)
    B, L = wav.shape
    T, pl, pr = tf_same_pad(L, n_dft, hop)
    x = np.pad(wav, ((0, 0), (pl, pr)))
    idx = np.arange(T)[:, None] * hop + np.arange(n_dft)[None, :]
    frames = x[:, idx] * np.asarray(window, dtype=dtype)[None, None, :]
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1)
    return (spec.real ** 2 + spec.imag ** 2).astype(dtype)


def power_spectrogram_dense(wav: np.ndarray, window: np.ndarray, n_dft=1024, hop=160, dtype=np.float64) -> np.ndarray:
    """Same as power_spectrogram but literally as the reference's dense DFT-kernel product (slow; small inputs)."""
    wav = np.asarray(wav, dtype=dtype)
    B, L = wav.shape
    T, pl, pr = tf_same_pad(L, n_dft, hop)
    x = np.pad(wav, ((0, 0), (pl, pr)))
    n = np.arange(n_dft)
    wk = np.arange(n_dft // 2 + 1) * 2 * np.pi / float(n_dft)
    real_k = (np.cos(wk[:, None] * n[None, :]) * window[None, :]).astype(dtype)      # backend.py:52-61
    imag_k = (-np.sin(wk[:, None] * n[None, :]) * window[None, :]).astype(dtype)
    idx = np.arange(T)[:, None] * hop + n[None, :]
    frames = x[:, idx]
    re = frames @ real_k.T
    im = frames @ imag_k.T
    return re ** 2 + im ** 2


def amplitude_to_decibel(p: np.ndarray, amin=1e-10, dynamic_range=80.0) -> np.ndarray:
    """backend_keras.amplitude_to_decibel (asr/models/layers/backend_keras.py:5-23): per-utterance max over all
    non-batch axes."""
    ln10 = np.log(np.asarray(10.0, dtype=np.float32)).astype(p.dtype)          # np.log(10).astype(floatx)
    log_spec = 10 * np.log(np.maximum(p, amin)) / ln10
    axes = tuple(range(1, p.ndim))
    log_spec = log_spec - log_spec.max(axis=axes, keepdims=True)
    return np.maximum(log_spec, -dynamic_range)


def melspectrogram(wav: np.ndarray, raw: Raw, dtype=np.float64) -> np.ndarray:
    """Melspectrogram.call (time_frequency.py:173-189): dB power spectrogram projected (linearly) on freq2mel.
    wav [B, L] -> [B, T, n_mels]."""
    p = power_spectrogram(wav, raw["fe.window"], dtype=dtype)
    db = amplitude_to_decibel(p)
    return db @ raw["fe.mel"].astype(dtype)


def chunk_melspectrogram(wav: np.ndarray, raw: Raw, dtype=np.float64) -> np.ndarray:
    """padding='valid' variant used by ChunkConformer (time_frequency.py:106-107; backend_keras.py:25-37):
    n_dft-1 zeros in front, log10(max(p,1e-10)) only."""
    wav = np.asarray(wav, dtype=dtype)
    n_dft, hop = 1024, 160
    x = np.pad(wav, ((0, 0), (n_dft - 1, 0)))
    T = (x.shape[1] - n_dft) // hop + 1
    idx = np.arange(T)[:, None] * hop + np.arange(n_dft)[None, :]
    frames = x[:, idx] * raw["fe.window"].astype(dtype)[None, None, :]
    spec = np.fft.rfft(frames, axis=-1)
    p = spec.real ** 2 + spec.imag ** 2
    ln10 = np.log(np.asarray(10.0, dtype=np.float32)).astype(p.dtype)
    db = np.log(np.maximum(p, 1e-10)) / ln10
    return db @ raw["fe.mel"].astype(dtype)


# ----------------------------------------------------------------------------------------------------- subsampling
def conv2d_same(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride=(2, 2)) -> np.ndarray:
    """tf.keras.layers.Conv2D(padding='same') NHWC x HWIO (conformer_blocks.py:76-85)."""
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = w.shape
    Ho, pt, pb = tf_same_pad(H, kh, stride[0])
    Wo, pl, pr = tf_same_pad(W, kw, stride[1])
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    out = np.zeros((B, Ho, Wo, Cout), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (Ho - 1) * stride[0] + 1:stride[0], j:j + (Wo - 1) * stride[1] + 1:stride[1], :]
            out += patch @ w[i, j].astype(x.dtype)
    return out + b.astype(x.dtype)


def conv_subsampling(mel: np.ndarray, raw: Raw) -> np.ndarray:
    """ConvSubsampling.call (conformer_blocks.py:90-96) + merge_two_last_dims (utils/tools.py:89-91).
    mel [B, T, n_mels] -> [B, T', D]."""
    x = mel[..., None]
    x = np.maximum(conv2d_same(x, raw["sub.conv1.w"], raw["sub.conv1.b"]), 0)
    x = np.maximum(conv2d_same(x, raw["sub.conv2.w"], raw["sub.conv2.b"]), 0)
    B, T2, F2, D = x.shape
    x = x.reshape(B, T2, F2 * D)
    return x @ raw["sub.lin.w"].astype(x.dtype) + raw["sub.lin.b"].astype(x.dtype)


# ----------------------------------------------------------------------------------------------------- conformer block
def ff_module(x, raw: Raw, p: str, fc_factor=0.5):
    """FFModule.call (conformer_blocks.py:126-134)."""
    dt = x.dtype
    h = layer_norm(x, raw[p + ".ln.g"].astype(dt), raw[p + ".ln.b"].astype(dt))
    h = swish(h @ raw[p + ".w1"].astype(dt) + raw[p + ".b1"].astype(dt))
    h = h @ raw[p + ".w2"].astype(dt) + raw[p + ".b2"].astype(dt)
    return x + fc_factor * h


def mhsa_module(x, raw: Raw, p: str, mask=None):
    """MHSAModule.call (conformer_blocks.py:163-170) + MultiHeadAttention.call (multihead_attention.py:151-188):
    no positional term, no q/k/v bias, one output bias.  mask [.., N, M] (1 = keep) as in :165-174."""
    dt = x.dtype
    xn = layer_norm(x, raw[p + ".ln.g"].astype(dt), raw[p + ".ln.b"].astype(dt))
    wq, wk, wv, wo = (raw[p + s].astype(dt) for s in (".wq", ".wk", ".wv", ".wo"))
    q = np.einsum("bni,hio->bnho", xn, wq)
    k = np.einsum("bmi,hio->bmho", xn, wk)
    v = np.einsum("bmi,hio->bmho", xn, wv)
    q = q / np.sqrt(np.asarray(wq.shape[-1], dtype=dt))
    logits = np.einsum("bnho,bmho->bhnm", q, k)
    if mask is not None:
        logits = logits + -10e9 * (1.0 - np.asarray(mask, dtype=dt))
    logits = logits - logits.max(-1, keepdims=True)
    e = np.exp(logits)
    coef = e / e.sum(-1, keepdims=True)
    o = np.einsum("bhnm,bmhi->bnhi", coef, v)
    out = np.einsum("bnhi,hio->bno", o, wo) + raw[p + ".bo"].astype(dt)
    return x + out


def depthwise_conv1d(x, w, pad_left: int, pad_right: int):
    """Depthwise part of tf.keras.layers.SeparableConv1D (cross-correlation).  x [B,T,C], w [K,C]."""
    K = w.shape[0]
    xp = np.pad(x, ((0, 0), (pad_left, pad_right), (0, 0)))
    T = x.shape[1]
    out = np.zeros_like(x)
    for j in range(K):
        out += xp[:, j:j + T, :] * w[j].astype(x.dtype)
    return out


def conv_module(x, raw: Raw, p: str, causal=False):
    """ConvModule.call (conformer_blocks.py:208-219): LN, pw1, GLU (first half * sigmoid(second half), :16-19),
    SeparableConv1D(k, 'same'), BatchNorm (eval, folded), swish, pw2, residual."""
    dt = x.dtype
    y = layer_norm(x, raw[p + ".ln.g"].astype(dt), raw[p + ".ln.b"].astype(dt))
    y = y @ raw[p + ".pw1.w"].astype(dt) + raw[p + ".pw1.b"].astype(dt)
    D = x.shape[-1]
    y = y[..., :D] * sigmoid(y[..., D:])
    K = raw[p + ".dw.w"].shape[0]
    if causal:
        pl, pr = K - 1, 0
    else:
        _, pl, pr = tf_same_pad(x.shape[1], K, 1)
    y = depthwise_conv1d(y, raw[p + ".dw.w"], pl, pr)
    y = y @ raw[p + ".pw.w"].astype(dt) + raw[p + ".pw.b"].astype(dt)
    y = y * raw[p + ".bn.scale"].astype(dt) + raw[p + ".bn.shift"].astype(dt)
    y = swish(y)
    y = y @ raw[p + ".pw2.w"].astype(dt) + raw[p + ".pw2.b"].astype(dt)
    return x + y


def conformer_block(x, raw: Raw, p: str, taps=None):
    """ConformerBlock.call (conformer_blocks.py:259-265)."""
    x = ff_module(x, raw, p + "ffn1")
    if taps is not None:
        taps[p + "ffn1"] = x
    x = mhsa_module(x, raw, p + "mhsa")
    if taps is not None:
        taps[p + "mhsa"] = x
    x = conv_module(x, raw, p + "conv")
    if taps is not None:
        taps[p + "conv"] = x
    x = ff_module(x, raw, p + "ffn2")
    if taps is not None:
        taps[p + "ffn2"] = x
    dt = x.dtype
    x = layer_norm(x, raw[p + "ln.g"].astype(dt), raw[p + "ln.b"].astype(dt))
    if taps is not None:
        taps[p + "out"] = x
    return x


# ----------------------------------------------------------------------------------------------------- models
def encoder_forward(wav: np.ndarray, raw: Raw, num_blocks: int, dtype=np.float64, taps=None) -> np.ndarray:
    """ConformerEncoder.call (conformer_blocks.py:343-356).  wav [B, L] -> [B, T', D]."""
    mel = melspectrogram(np.asarray(wav), raw, dtype=dtype)
    if taps is not None:
        taps["mel"] = mel
    x = conv_subsampling(mel, raw)
    if taps is not None:
        taps["sub"] = x
    for i in range(num_blocks):
        x = conformer_block(x, raw, f"enc.{i}.", taps)
    return x


def ctc_forward(enc: np.ndarray, raw: Raw, num_blocks: int = 1, dtype=np.float64, taps=None) -> np.ndarray:
    """CTCDecoder.call (conformer_blocks.py:419-424).  enc [B, T', D] -> logits [B, T', V]."""
    x = np.asarray(enc, dtype=dtype)
    x = x @ raw["ctc.proj.w"].astype(dtype) + raw["ctc.proj.b"].astype(dtype)
    for i in range(num_blocks):
        x = conformer_block(x, raw, f"ctc.blk{i}.", taps)
    if taps is not None:
        taps["ctc.hidden"] = x
    return x @ raw["ctc.fc.w"].astype(dtype) + raw["ctc.fc.b"].astype(dtype)


def streaming_encoder_forward(wav: np.ndarray, raw: Raw, num_blocks: int, chunk: int = 8000, dtype=np.float64):
    """StreamingConformerEncoder.call (conformer_blocks.py:574-594): split into independent `chunk`-sample pieces
    (zero-padded to a multiple of chunk), encode each alone, concatenate in time.  wav [B, L] -> [B, n*13, D]."""
    wav = np.asarray(wav)
    B, L = wav.shape
    n = -(-L // chunk)
    x = np.pad(wav, ((0, 0), (0, n * chunk - L))).reshape(B * n, chunk)
    enc = encoder_forward(x, raw, num_blocks, dtype=dtype)
    return enc.reshape(B, n * enc.shape[1], enc.shape[2])


# ----------------------------------------------------------------------------------------------------- translator
def positional_encoding(max_len: int, size: int, dtype=np.float64) -> np.ndarray:
    """positional_encoding (asr/models/layers/positional_encoding.py:19-36): sin on the even columns, cos on the odd columns,
    both with the exponent 2 * (index // 2) / size.  -> [max_len, size]."""
    pos = np.arange(max_len, dtype=np.float32)[:, None]
    index = np.arange(size, dtype=np.float32)[None, :]
    pe = pos * (np.float32(1.0) / np.power(np.float32(10000.0), (2 * (index // 2)) / np.float32(size)))
    out = np.zeros((max_len, size), dtype=np.float32)
    out[:, 0::2] = np.sin(pe[:, 0::2])
    out[:, 1::2] = np.cos(pe[:, 1::2])
    return out.astype(dtype)


def rmhsa_module(x, enc, raw: Raw, p: str):
    """RMHSAModule.call (conformer_blocks.py:454-463): queries = LN(x + positional encoding), keys = values = the encoder states
    (no LayerNorm, no positional term on them), residual on x (without the positional encoding)."""
    dt = x.dtype
    q_in = layer_norm(x + positional_encoding(x.shape[1], x.shape[2], dt)[None], raw[p + ".ln.g"].astype(dt), raw[p + ".ln.b"].astype(dt))
    wq, wk, wv, wo = (raw[p + s].astype(dt) for s in (".wq", ".wk", ".wv", ".wo"))
    q = np.einsum("bni,hio->bnho", q_in, wq) / np.sqrt(np.asarray(wq.shape[-1], dtype=dt))
    k = np.einsum("bmi,hio->bmho", enc, wk)
    v = np.einsum("bmi,hio->bmho", enc, wv)
    logits = np.einsum("bnho,bmho->bhnm", q, k)
    logits = logits - logits.max(-1, keepdims=True)
    e = np.exp(logits)
    coef = e / e.sum(-1, keepdims=True)
    o = np.einsum("bhnm,bmhi->bnhi", coef, v)
    return x + np.einsum("bnhi,hio->bno", o, wo) + raw[p + ".bo"].astype(dt)


def rblock(x, enc, raw: Raw, p: str):
    """RBlock.call (conformer_blocks.py:496-502)."""
    x = ff_module(x, raw, p + "ffn1")
    x = rmhsa_module(x, enc, raw, p + "mhsa")
    x = conv_module(x, raw, p + "conv")
    x = ff_module(x, raw, p + "ffn2")
    dt = x.dtype
    return layer_norm(x, raw[p + "ln.g"].astype(dt), raw[p + "ln.b"].astype(dt))


def translator_forward(ids: np.ndarray, enc: np.ndarray, raw: Raw, num_blocks: int, dtype=np.float64) -> np.ndarray:
    """Translator.call (conformer_blocks.py:546-552): embedding -> N x RBlock(x, enc) -> Dense.  ids [B, U] int, enc [B, T', D]
    -> [B, U, tar_classes]."""
    x = np.asarray(raw["tr.emb"], dtype=dtype)[np.asarray(ids)]
    enc = np.asarray(enc, dtype=dtype)
    for i in range(num_blocks):
        x = rblock(x, enc, raw, f"tr.{i}.")
    return x @ raw["tr.fc.w"].astype(dtype) + raw["tr.fc.b"].astype(dtype)
