"""TEST INFRASTRUCTURE ONLY (oracle) -- NumPy restatement of the reference's ChunkConformer:
the causal chunk-streaming model with state caches, offline (`predict`) and streaming (`picker_stream_predict`,
`feature_pick`, `decoder_stream_predict`) paths, each function citing the reference lines it follows
(asr/models/chunk_conformer_blocks.py unless stated otherwise).

Parity status: **UNPINNED**.  The reference ships no ChunkConformer weights and TensorFlow cannot be imported in this image,
so neither golden vectors nor a live comparison exist; weights here are seeded random (`random_chunk_model`).  What IS
checked (tests/test_chunk_oracle.py) is the reference's own consistency criterion (test_chunk_asr.py:57,123,139): the
streaming path reproduces the offline path frame for frame wherever no look-ahead is involved (front end, encoder and
picker use win_back = 0), and the decoder's "valid" frames reproduce the offline decoder.  No CUDA path exists for this row
yet; this file is the specification the next round builds against.

Arithmetic in float64 unless `dtype` says otherwise.  Third-party semantics restated from their documentation:
tf.keras.layers.MultiHeadAttention (biases on q/k/v/o, queries scaled by 1/sqrt(key_dim), boolean attention_mask True = attend),
Conv1D / SeparableConv1D padding='causal' (kernel_size-1 zeros on the left), LayerNormalization / BatchNormalization eps 1e-3.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import conformer_ref as cr

Raw = Dict[str, np.ndarray]

CFG = dict(dmodel=144, n_mels=80, chunk_num=16, reduction=4, hop=160, n_dft=1024,
           enc_blocks=15, heads=4, head_size=36, kernel=32, win_front=36,
           picker_blocks=1, picker_back=0, phone_classes=277,
           helper_blocks=2, dec_blocks=1, dec_back=8, txt_classes=9171)          # asr/configs/chunk_conformerS.yml


# ----------------------------------------------------------------------------------------------------- weights
def random_chunk_model(seed: int, fe_raw: Raw, cfg: dict = CFG, scale: float = 1.0) -> Raw:
    """Seeded random weights with the reference's tensor shapes.  fe_raw supplies 'fe.window' [1024] and 'fe.mel' [513, 80]
    (the Hann window / mel matrix baked into the shipped ONNX graphs; the ChunkConformer front end builds the same ones,
    time_frequency.py:152-160)."""
    rng = np.random.default_rng(seed)
    D, H, dh, K = cfg["dmodel"], cfg["heads"], cfg["head_size"], cfg["kernel"]
    raw: Raw = {"fe.window": np.asarray(fe_raw["fe.window"], np.float64), "fe.mel": np.asarray(fe_raw["fe.mel"], np.float64)}

    def dense(name, i, o):
        raw[name + ".w"] = rng.standard_normal((i, o)) * scale / np.sqrt(i)
        raw[name + ".b"] = rng.standard_normal(o) * 0.1

    def ln(name):
        raw[name + ".g"] = 1.0 + 0.1 * rng.standard_normal(D)
        raw[name + ".b"] = 0.1 * rng.standard_normal(D)

    def block(p):
        for f in ("ffn1", "ffn2"):
            ln(p + f + ".ln")
            raw[p + f + ".w1"] = rng.standard_normal((D, 4 * D)) * scale / np.sqrt(D)
            raw[p + f + ".b1"] = rng.standard_normal(4 * D) * 0.1
            raw[p + f + ".w2"] = rng.standard_normal((4 * D, D)) * scale / np.sqrt(4 * D)
            raw[p + f + ".b2"] = rng.standard_normal(D) * 0.1
        ln(p + "mhsa.ln")
        for n in ("q", "k", "v"):
            raw[p + f"mhsa.w{n}"] = rng.standard_normal((D, H, dh)) * scale / np.sqrt(D)
            raw[p + f"mhsa.b{n}"] = rng.standard_normal((H, dh)) * 0.1
        raw[p + "mhsa.wo"] = rng.standard_normal((H, dh, D)) * scale / np.sqrt(H * dh)
        raw[p + "mhsa.bo"] = rng.standard_normal(D) * 0.1
        ln(p + "conv.ln")
        raw[p + "conv.pw1.w"] = rng.standard_normal((D, 2 * D)) * scale / np.sqrt(D)
        raw[p + "conv.pw1.b"] = rng.standard_normal(2 * D) * 0.1
        raw[p + "conv.dw.w"] = rng.standard_normal((K, D)) * scale / np.sqrt(K)
        raw[p + "conv.pw.w"] = rng.standard_normal((D, 2 * D)) * scale / np.sqrt(D)
        raw[p + "conv.pw.b"] = rng.standard_normal(2 * D) * 0.1
        raw[p + "conv.bn.scale"] = 1.0 + 0.1 * rng.standard_normal(2 * D)      # gamma / sqrt(var + 1e-3), eval mode
        raw[p + "conv.bn.shift"] = 0.1 * rng.standard_normal(2 * D)
        raw[p + "conv.pw2.w"] = rng.standard_normal((2 * D, D)) * scale / np.sqrt(2 * D)
        raw[p + "conv.pw2.b"] = rng.standard_normal(D) * 0.1
        ln(p + "ln")

    raw["sub.conv1.w"] = rng.standard_normal((3, 3, 1, D)) * scale / 3.0
    raw["sub.conv1.b"] = rng.standard_normal(D) * 0.1
    raw["sub.conv2.w"] = rng.standard_normal((3, 3, D, D)) * scale / np.sqrt(9 * D)
    raw["sub.conv2.b"] = rng.standard_normal(D) * 0.1
    F2 = ((cfg["n_mels"] + 4 - 3) // 2 + 1 - 3) // 2 + 1                        # 80 -> pad 84 -> 41 -> 20
    dense("sub.lin", F2 * D, D)
    for i in range(cfg["enc_blocks"]):
        block(f"enc.{i}.")
    dense("picker.proj", D, D)
    for i in range(cfg["picker_blocks"]):
        block(f"picker.{i}.")
    dense("picker.fc", D, cfg["phone_classes"])
    for i in range(cfg["helper_blocks"]):
        block(f"helper.{i}.")
    dense("dec.proj", D, D)
    for i in range(cfg["dec_blocks"]):
        block(f"dec.{i}.")
    dense("dec.fc", D, cfg["txt_classes"])
    return raw


# ----------------------------------------------------------------------------------------------------- front end
def mel_valid(wav: np.ndarray, raw: Raw) -> np.ndarray:
    """Melspectrogram(padding='valid') (time_frequency.py:100-122,173-189 + backend_keras.py:25-37): n_dft-1 zeros in FRONT of
    whatever it is given (:106-107), stride-160 'valid' frames, |X|^2, log10(max(p, 1e-10)) WITHOUT per-utterance max, mel matrix.
    wav [B, L] -> [B, floor((L - 1) / 160) + 1, 80]; frame t ends at sample 160 t."""
    return cr.chunk_melspectrogram(wav, raw)


def conv2d_valid(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride=(2, 2)) -> np.ndarray:
    """Conv2D(3x3, strides, padding='valid') + ReLU.  x [B, T, F, Cin], w [3, 3, Cin, Cout] (Keras HWIO)."""
    B, T, F, Cin = x.shape
    To, Fo = (T - 3) // stride[0] + 1, (F - 3) // stride[1] + 1
    out = np.zeros((B, max(To, 0), Fo, w.shape[-1]), x.dtype)
    for kh in range(3):
        for kw in range(3):
            patch = x[:, kh:kh + stride[0] * (To - 1) + 1:stride[0], kw:kw + stride[1] * (Fo - 1) + 1:stride[1], :]
            out += np.einsum("btfc,co->btfo", patch, w[kh, kw].astype(x.dtype))
    return np.maximum(out + b.astype(x.dtype), 0)


def subsample_tail(x: np.ndarray, raw: Raw) -> np.ndarray:
    """conv1 (stride (reduction/2, 2) = (2, 2)) -> conv2 (2, 2) -> merge_two_last_dims -> Dense (:56-71, :74-96).  x is the mel map
    AFTER the time padding / cache concatenation, [B, T, 80]; the frequency padding [2, 2] is applied here."""
    x = np.pad(x, ((0, 0), (0, 0), (2, 2)))[..., None]
    x = conv2d_valid(x, raw["sub.conv1.w"], raw["sub.conv1.b"])
    x = conv2d_valid(x, raw["sub.conv2.w"], raw["sub.conv2.b"])
    B, T2, F2, D = x.shape
    return x.reshape(B, T2, F2 * D) @ raw["sub.lin.w"] + raw["sub.lin.b"]


def front_call(wav: np.ndarray, raw: Raw) -> np.ndarray:
    """ChunkConformerFront.call (:447-453) + ConvSubsampling.call (:56-71): 4 zero mel frames in front ('valid' branch)."""
    mel = mel_valid(wav, raw)
    return subsample_tail(np.pad(mel, ((0, 0), (4, 0), (0, 0))), raw)


def front_init_caches(B: int, cfg: dict = CFG):
    """ChunkConformerFront.init_caches (:442-444): empty wav cache, sub_length = chunk_num / reduction = 4 zero mel frames."""
    return np.zeros((B, 0)), np.zeros((B, cfg["chunk_num"] // cfg["reduction"], cfg["n_mels"]))


def front_stream_call(wav_chunk: np.ndarray, wav_cache: np.ndarray, sub_cache: np.ndarray, raw: Raw, cfg: dict = CFG):
    """ChunkConformerFront.stream_call (:455-466) + ConvSubsampling.stream_call (:74-91): mel of [cache | chunk], keep the last
    chunk_num frames, prepend the mel cache, subsample, keep the last T = chunk_num / reduction outputs; the new wav cache is
    the last chunk_num * 160 samples, the new mel cache the last 4 mel frames."""
    buf = np.concatenate([wav_cache, wav_chunk], axis=1)
    mel = mel_valid(buf, raw)[:, -cfg["chunk_num"]:]
    cat = np.concatenate([sub_cache, mel], axis=1)
    out = subsample_tail(cat, raw)[:, -(cfg["chunk_num"] // cfg["reduction"]):]
    return out, buf[:, -cfg["chunk_num"] * cfg["hop"]:], cat[:, -(cfg["chunk_num"] // cfg["reduction"]):]


# ----------------------------------------------------------------------------------------------------- blocks
def chunk_mask(n: int, win_front: int, win_back: int) -> np.ndarray:
    """ChunkMHSAModule._compute_chunk_mask (:158-176): [n, n] boolean, True = attend."""
    idx = np.arange(n)[:, None]
    col = np.arange(n)[None, :]
    low = np.maximum(idx - win_front, 0)
    high = np.clip(idx + win_back, 0, n)
    low = low - np.maximum(low - n + win_back, 0)
    high = high + np.maximum(win_back - high, 0)
    return ~((col < low) | (col > high))


def keras_mha(q_in: np.ndarray, kv_in: np.ndarray, raw: Raw, p: str, mask: np.ndarray) -> np.ndarray:
    """tf.keras.layers.MultiHeadAttention(num_heads, key_dim=head_size)(query, value, attention_mask) (:145): projections with
    bias, q / sqrt(key_dim), masked softmax (masked logits get -1e9, Keras' Softmax layer), output projection with bias."""
    wq, wk, wv, wo = (raw[p + s] for s in (".wq", ".wk", ".wv", ".wo"))
    q = np.einsum("bnd,dhk->bnhk", q_in, wq) + raw[p + ".bq"]
    k = np.einsum("bmd,dhk->bmhk", kv_in, wk) + raw[p + ".bk"]
    v = np.einsum("bmd,dhk->bmhk", kv_in, wv) + raw[p + ".bv"]
    q = q / np.sqrt(float(wq.shape[-1]))
    s = np.einsum("bnhk,bmhk->bhnm", q, k)
    s = np.where(mask[None, None], s, -1e9)
    s = s - s.max(-1, keepdims=True)
    e = np.exp(s)
    a = e / e.sum(-1, keepdims=True)
    o = np.einsum("bhnm,bmhk->bnhk", a, v)
    return np.einsum("bnhk,hkd->bnd", o, wo) + raw[p + ".bo"]


def mhsa_call(x, raw, p, win_front, win_back):
    """ChunkMHSAModule.call (:192-200)."""
    xn = cr.layer_norm(x, raw[p + ".ln.g"], raw[p + ".ln.b"])
    return x + keras_mha(xn, xn, raw, p, chunk_mask(x.shape[1], win_front, win_back))


def mhsa_stream_call(x, cache, raw, p, win_front, win_back):
    """ChunkMHSAModule.stream_call (:202-216): LayerNorm over [cache | x], mask over the concatenation, queries = the last T rows."""
    T = x.shape[1]
    cat = np.concatenate([cache, x], axis=1)
    xn = cr.layer_norm(cat, raw[p + ".ln.g"], raw[p + ".ln.b"])
    mask = chunk_mask(cat.shape[1], win_front, win_back)[-T:]
    return x + keras_mha(xn[:, -T:], xn, raw, p, mask), cat


def conv_stream_call(x, cache, raw, p):
    """ChunkConvModule.stream_call (:294-311): the whole module over [cache | x] with 'causal' padding, last T rows, residual on x."""
    T = x.shape[1]
    cat = np.concatenate([cache, x], axis=1)
    y = cr.conv_module(cat, raw, p, causal=True) - cat            # module output without its own residual
    return x + y[:, -T:], cat


def block_call(x, raw, p, win_front, win_back):
    """ChunkConformerBlock.call (:374-380); blocks are built with padding='causal' (:334, :484)."""
    x = cr.ff_module(x, raw, p + "ffn1")
    x = mhsa_call(x, raw, p + "mhsa", win_front, win_back)
    x = cr.conv_module(x, raw, p + "conv", causal=True)
    x = cr.ff_module(x, raw, p + "ffn2")
    return cr.layer_norm(x, raw[p + "ln.g"], raw[p + "ln.b"])


def block_stream_call(x, mha_cache, cnn_cache, raw, p, win_front, win_back):
    """ChunkConformerBlock.stream_call (:382-389)."""
    x = cr.ff_module(x, raw, p + "ffn1")
    x, mha_new = mhsa_stream_call(x, mha_cache, raw, p + "mhsa", win_front, win_back)
    x, cnn_new = conv_stream_call(x, cnn_cache, raw, p + "conv")
    x = cr.ff_module(x, raw, p + "ffn2")
    return cr.layer_norm(x, raw[p + "ln.g"], raw[p + "ln.b"]), mha_new, cnn_new


def stack_init_caches(B: int, n_blocks: int, cfg: dict = CFG):
    """init_caches of the encoder / picker / decoder / helper (:511-521, :605-614, :712-721): empty caches per block."""
    z = [np.zeros((B, 0, cfg["dmodel"])) for _ in range(n_blocks)]
    return list(z), [c.copy() for c in z]


def stack_stream_call(x, mha_caches, cnn_caches, raw, prefix, n_blocks, win_front, win_back, kernel):
    """The common body of ChunkConformerEncoder / ChunkCTCDecoder / ContextHelper.stream_call (:532-563, :626-658, :735-758):
    every block streams; with look-ahead (win_back > 0) the last win_back frames are 'unvalid' and are dropped from the caches
    before the caches are cut to win_front / kernel_size rows.  Returns (all block outputs, new_mha, new_cnn)."""
    new_mha, new_cnn = [], []
    for i in range(n_blocks):
        x, m, c = block_stream_call(x, mha_caches[i], cnn_caches[i], raw, f"{prefix}{i}.", win_front, win_back)
        new_mha.append(m)
        new_cnn.append(c)
    if win_back != 0:
        new_mha = [m[:, :-win_back] for m in new_mha]
        new_cnn = [c[:, :-win_back] for c in new_cnn]
    new_mha = [m[:, -win_front:] for m in new_mha]
    new_cnn = [c[:, -kernel:] for c in new_cnn]
    return x, new_mha, new_cnn


# ----------------------------------------------------------------------------------------------------- model
def encoder_call(x, raw, cfg: dict = CFG):
    for i in range(cfg["enc_blocks"]):
        x = block_call(x, raw, f"enc.{i}.", cfg["win_front"], 0)
    return x


def ctc_decoder_call(x, raw, name, n_blocks, win_back, cfg: dict = CFG):
    """ChunkCTCDecoder.call (:617-622): project, blocks, fc -> (ctc logits, hidden)."""
    x = x @ raw[name + ".proj.w"] + raw[name + ".proj.b"]
    for i in range(n_blocks):
        x = block_call(x, raw, f"{name}.{i}.", cfg["win_front"], win_back)
    return x @ raw[name + ".fc.w"] + raw[name + ".fc.b"], x


def ctc_decoder_stream_call(x, mha_caches, cnn_caches, raw, name, n_blocks, win_back, cfg: dict = CFG):
    """ChunkCTCDecoder.stream_call (:626-658): (valid logits, valid hidden, caches, unvalid logits)."""
    x = x @ raw[name + ".proj.w"] + raw[name + ".proj.b"]
    h, m, c = stack_stream_call(x, mha_caches, cnn_caches, raw, name + ".", n_blocks, cfg["win_front"], win_back, cfg["kernel"])
    logits = h @ raw[name + ".fc.w"] + raw[name + ".fc.b"]
    if win_back != 0:
        return logits[:, :-win_back], h[:, :-win_back], m, c, logits[:, -win_back:]
    return logits, h, m, c, np.zeros_like(logits)


def helper_call(x, raw, cfg: dict = CFG):
    """ContextHelper.call (:723-728)."""
    for i in range(cfg["helper_blocks"]):
        x = block_call(x, raw, f"helper.{i}.", cfg["win_front"], 0)
    return x


def feature_pick(hidden: np.ndarray, ctc: np.ndarray, blank: int) -> Tuple[np.ndarray, np.ndarray]:
    """ChunkConformer.feature_pick (:913-999): keep the frames whose phone argmax is not blank (= num_classes - 1), compact them
    to the front, zero-pad every row to the longest."""
    keep = ctc.argmax(-1) != blank
    n = int(keep.sum(1).max()) if keep.size else 0
    B = hidden.shape[0]
    f = np.zeros((B, n, hidden.shape[-1]), hidden.dtype)
    c = np.zeros((B, n, ctc.shape[-1]), ctc.dtype)
    for b in range(B):
        k = int(keep[b].sum())
        f[b, :k] = hidden[b, keep[b]]
        c[b, :k] = ctc[b, keep[b]]
    return f, c


def predict(wav: np.ndarray, raw: Raw, cfg: dict = CFG) -> np.ndarray:
    """ChunkConformer.predict (:798-805): offline text logits [B, U, txt_classes]."""
    x = encoder_call(front_call(wav, raw), raw, cfg)
    phone, hidden = ctc_decoder_call(x, raw, "picker", cfg["picker_blocks"], cfg["picker_back"], cfg)
    picked, _ = feature_pick(hidden, phone, cfg["phone_classes"] - 1)
    out, _ = ctc_decoder_call(helper_call(picked, raw, cfg), raw, "dec", cfg["dec_blocks"], cfg["dec_back"], cfg)
    return out


def init_picker_caches(B: int, cfg: dict = CFG):
    """ChunkConformer.init_picker_caches (:777-786)."""
    wav_c, sub_c = front_init_caches(B, cfg)
    em, ec = stack_init_caches(B, cfg["enc_blocks"], cfg)
    pm, pc = stack_init_caches(B, cfg["picker_blocks"], cfg)
    return [wav_c, sub_c, em, ec, pm, pc, np.zeros((B, 0, cfg["dmodel"]))]


def init_decoder_caches(B: int, cfg: dict = CFG):
    """ChunkConformer.init_decoder_caches (:788-792)."""
    hm, hc = stack_init_caches(B, cfg["helper_blocks"], cfg)
    dm, dc = stack_init_caches(B, cfg["dec_blocks"], cfg)
    return [hm, hc, dm, dc, np.zeros((B, 0, cfg["dmodel"]))]


def picker_stream_predict(wav_chunk: np.ndarray, caches: list, raw: Raw, cfg: dict = CFG):
    """ChunkConformer.picker_stream_predict (:807-824): (valid phone logits, unvalid phone logits, valid hidden, caches)."""
    wav_c, sub_c, em, ec, pm, pc, dec_inp = caches
    x, wav_c, sub_c = front_stream_call(wav_chunk, wav_c, sub_c, raw, cfg)
    x, em, ec = stack_stream_call(x, em, ec, raw, "enc.", cfg["enc_blocks"], cfg["win_front"], 0, cfg["kernel"])
    dec_inp = np.concatenate([dec_inp, x], axis=1)
    valid, hidden, pm, pc, unvalid = ctc_decoder_stream_call(dec_inp, pm, pc, raw, "picker", cfg["picker_blocks"], cfg["picker_back"], cfg)
    dec_inp = dec_inp[:, valid.shape[1]:]
    return valid, unvalid, hidden, [wav_c, sub_c, em, ec, pm, pc, dec_inp]


def decoder_stream_predict(features: np.ndarray, caches: list, raw: Raw, cfg: dict = CFG):
    """ChunkConformer.decoder_stream_predict (:826-837): (valid text logits, unvalid text logits, caches)."""
    hm, hc, dm, dc, dec_inp = caches
    x, hm, hc = stack_stream_call(features, hm, hc, raw, "helper.", cfg["helper_blocks"], cfg["win_front"], 0, cfg["kernel"])
    dec_inp = np.concatenate([dec_inp, x], axis=1)
    valid, _, dm, dc, unvalid = ctc_decoder_stream_call(dec_inp, dm, dc, raw, "dec", cfg["dec_blocks"], cfg["dec_back"], cfg)
    dec_inp = dec_inp[:, valid.shape[1]:]
    return valid, unvalid, [hm, hc, dm, dc, dec_inp]


def stream_utterance(wav: np.ndarray, raw: Raw, cfg: dict = CFG):
    """The driver loop of test_chunk_asr.py:47-93 (B = 1): chunk the waveform into chunk_num * 160 samples, run the picker step,
    compact, run the decoder step when something was picked.  Returns (valid text logits, last unvalid text logits,
    picked phone logits)."""
    step = cfg["chunk_num"] * cfg["hop"]
    c1, c2 = init_picker_caches(1, cfg), init_decoder_caches(1, cfg)
    txt = np.zeros((1, 0, cfg["txt_classes"]))
    phones = np.zeros((1, 0, cfg["phone_classes"]))
    unvalid_txt = np.zeros((1, 0, cfg["txt_classes"]))
    for s in range(0, wav.shape[1], step):
        v_ph, _, v_hid, c1 = picker_stream_predict(wav[:, s:s + step], c1, raw, cfg)
        if v_ph.shape[1] == 0:
            continue
        feats, picked = feature_pick(v_hid, v_ph, cfg["phone_classes"] - 1)
        if feats.shape[1] != 0:
            v_txt, unvalid_txt, c2 = decoder_stream_predict(feats, c2, raw, cfg)
            txt = np.concatenate([txt, v_txt], axis=1)
            phones = np.concatenate([phones, picked], axis=1)
    return txt, unvalid_txt, phones
