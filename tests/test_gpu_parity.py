"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI (ctypes binding in
tensorflowasr_b200/engine.py), against the CPU oracle on the same seeded inputs and against golden vectors generated
from the reference itself.

Tolerances (stated per north_star "within a stated fp32 tolerance"):
  * fp32 mode (CUDA-core GEMMs): |enc - oracle| <= 2e-4, |logits - oracle| <= 2e-3 (values up to ~30)
  * tf32 mode (tcgen05 tensor cores, fp32 accumulate, every operand rounded to nearest tf32 by its producer):
    |enc - oracle| <= 8e-3; logits rms <= 1e-2 and max <= 0.25 (the max is an outlier bound, see tests/conftest.py)
  * mel (always fp32): 5e-3 dB on a [-80, 0] scale
  * token ids (greedy, beam): bit-exact
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_IDS, TOL_ENC, TOL_LOGITS_MAX, check_logits

pytestmark = pytest.mark.gpu

TOL = {p: dict(enc=TOL_ENC[p], logits=TOL_LOGITS_MAX[p]) for p in (0, 1)}


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


@pytest.fixture(scope="module", params=[1, 0], ids=["fp32", "tf32"])
def eng(request, offline_weights):
    from tensorflowasr_b200 import engine as E
    ge, re_, gc, rc = offline_weights
    e = E.Engine(ge, re_, gc, rc, precision=request.param, use_cuda_graph=True)
    e.precision = request.param
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng32(offline_weights):
    from tensorflowasr_b200 import engine as E
    ge, re_, gc, rc = offline_weights
    e = E.Engine(ge, re_, gc, rc, precision=1, use_cuda_graph=False)
    yield e
    e.close()


def test_native_library_is_loaded(eng32):
    """The product path is the in-tree CUDA library, not a framework fallback."""
    maps = open("/proc/self/maps").read()
    assert "libb200asr.so" in maps
    assert eng32.launch_count == 0


@pytest.mark.parametrize("L", [1, 159, 160, 1023, 1024, 16000, 67263])
def test_mel_parity_ragged_lengths(eng32, offline_weights, ref_wav, L):
    from oracle import conformer_ref as cr
    _, re_, _, _ = offline_weights
    x = ref_wav[None, 5000:5000 + L] if L < 60000 else ref_wav[None, :L]
    got = eng32.mel(x).cpu().numpy()
    ref = cr.melspectrogram(x, re_)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, atol=5e-3)


def test_mel_vs_reference_golden(eng32, golden, ref_wav):
    np.testing.assert_allclose(eng32.mel(ref_wav[None]).cpu().numpy()[0], golden["wav_mel"], atol=5e-3)


def test_mel_silence_and_scale_invariance(eng32, ref_wav):
    z = eng32.mel(np.zeros((1, 4000), np.float32)).cpu().numpy()
    assert np.isfinite(z).all() and np.allclose(z, 0.0)          # all bins at the 1e-10 floor == the max -> 0 dB
    a = eng32.mel(ref_wav[None, :20000]).cpu().numpy()
    b = eng32.mel(0.25 * ref_wav[None, :20000]).cpu().numpy()
    np.testing.assert_allclose(a, b, atol=2e-2)                   # per-utterance max normalisation


def test_encoder_logits_ids_on_reference_wav(eng, offline_weights, golden, ref_wav, torch_mod):
    from oracle import conformer_ref as cr
    ge, re_, gc, rc = offline_weights
    tol = TOL[eng.precision]
    enc = eng.encode(ref_wav[None])
    assert tuple(enc.shape) == (1, 106, 144)
    enc_ref = cr.encoder_forward(ref_wav[None], re_, ge.num_blocks)
    np.testing.assert_allclose(enc.cpu().numpy(), enc_ref, atol=tol["enc"])
    np.testing.assert_allclose(enc.cpu().numpy()[0], golden["wav_enc"], atol=tol["enc"] + 2e-4)
    logits = eng.ctc_logits(enc)
    lg_ref = cr.ctc_forward(enc_ref, rc, gc.num_blocks)
    check_logits(logits.cpu().numpy(), lg_ref, eng.precision, "reference wav vs oracle:")
    np.testing.assert_allclose(logits.cpu().numpy()[0][golden["wav_logit_frames"]], golden["wav_logits"], atol=tol["logits"] + 2e-3)
    assert (logits.cpu().numpy()[0].argmax(-1) == golden["wav_argmax"]).all()    # every frame's argmax = the reference's
    ids, lens = eng.ctc_greedy(logits)
    assert ids[0, :int(lens[0])].tolist() == GOLDEN_IDS          # bit-identical greedy ids (north_star)
    assert (ids[0, int(lens[0]):] == -1).all()
    ids2, lens2 = eng.recognize(ref_wav[None])
    assert ids2[0, :int(lens2[0])].tolist() == GOLDEN_IDS
    hi, hl = eng.recognize_host(torch_mod.from_numpy(ref_wav[None]).pin_memory())
    assert hi[0, :int(hl[0])].tolist() == GOLDEN_IDS


def test_noise_batch_vs_oracle_and_golden(eng, offline_weights, golden, noise_2x2s):
    from oracle import conformer_ref as cr
    ge, re_, gc, rc = offline_weights
    tol = TOL[eng.precision]
    enc = eng.encode(noise_2x2s).cpu().numpy()
    np.testing.assert_allclose(enc, cr.encoder_forward(noise_2x2s, re_, ge.num_blocks), atol=tol["enc"])
    np.testing.assert_allclose(enc, golden["noise_enc"], atol=tol["enc"] + 3e-4)
    ids, lens = eng.recognize(noise_2x2s)
    logits = eng.ctc_logits(eng.encode(noise_2x2s)).cpu().numpy()
    assert (logits.argmax(-1) == golden["noise_argmax"]).all()       # both precision modes: the reference's per-frame argmax


@pytest.mark.parametrize("L", [640, 4000, 8000, 12345, 31999])
def test_ragged_lengths_vs_oracle(eng, offline_weights, ref_wav, L):
    """Odd lengths exercise every 'same' padding branch (odd mel / conv frame counts)."""
    from oracle import conformer_ref as cr, ctc_ref
    ge, re_, gc, rc = offline_weights
    tol = TOL[eng.precision]
    x = ref_wav[None, 10000:10000 + L]
    enc_ref = cr.encoder_forward(x, re_, ge.num_blocks)
    enc = eng.encode(x)
    assert tuple(enc.shape) == enc_ref.shape
    np.testing.assert_allclose(enc.cpu().numpy(), enc_ref, atol=tol["enc"])
    lg_ref = cr.ctc_forward(enc_ref, rc, 1)
    np.testing.assert_allclose(eng.ctc_logits(enc).cpu().numpy(), lg_ref, atol=tol["logits"])
    ids, lens = eng.recognize(x)
    assert ids[0, :int(lens[0])].tolist() == ctc_ref.greedy_decode(lg_ref[0], 1331)


def test_batch_consistency_and_permutation(eng, ref_wav):
    """Equal-length batches: every row equals the single-utterance result; permuting rows permutes ids."""
    x = np.stack([ref_wav[:40000], ref_wav[20000:60000], ref_wav[:40000], ref_wav[27263:67263]])
    ids, lens = eng.recognize(x)
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    for i in range(4):
        si, sl = eng.recognize(x[i:i + 1])
        assert sl[0].item() == lens[i] and (si[0].cpu().numpy() == ids[i]).all()
    assert (ids[0] == ids[2]).all()
    pi, pl = eng.recognize(x[[3, 1, 0, 2]])
    assert (pi.cpu().numpy() == ids[[3, 1, 0, 2]]).all()


def test_full_size_batch_properties(eng, ref_wav, torch_mod):
    """BASELINE config 2 shape (32 x 10 s): finite outputs, tiled-speech rows decode identically to a single row, and the
    host-buffer entry point agrees with the device-buffer one."""
    L = 160000
    speech = np.tile(ref_wav, 3)[:L]
    rng = np.random.default_rng(1234)
    x = np.clip(rng.standard_normal((32, L)).astype(np.float32) * 0.1, -1, 1)
    x[::4] = speech
    xs = torch_mod.from_numpy(x).cuda()
    ids, lens = eng.recognize(xs)
    enc = eng.encode(xs)
    assert torch_mod.isfinite(enc).all()
    assert tuple(enc.shape) == (32, 250, 144)
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    one_ids, one_len = eng.recognize(speech[None])
    for r in range(0, 32, 4):
        assert lens[r] == one_len[0].item() and (ids[r] == one_ids[0].cpu().numpy()).all()
    assert lens[0] >= 30                                            # ~3 repetitions of a 13-token utterance
    hi, hl = eng.recognize_host(torch_mod.from_numpy(x).pin_memory())
    assert (hi.numpy() == ids).all() and (hl.numpy() == lens).all()


# ------------------------------------------------------------------------------------------------- CTC decoders
def test_greedy_kernel_vs_oracle(eng32, torch_mod):
    from oracle import ctc_ref
    rng = np.random.default_rng(5)
    B, T, V = 5, 37, 23
    logits = rng.standard_normal((B, T, V)).astype(np.float32) * 2
    logits[1, :, V - 1] = 50.0                                      # all blank
    logits[2] = 0.0                                                 # ties everywhere -> class 0 every frame
    logits[3, :, 4] = 9.0                                           # one long repeat
    lengths = np.array([37, 37, 37, 20, 0], np.int32)
    ids, lens = eng32.ctc_greedy(logits, lengths=lengths)
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    for b in range(B):
        ref = ctc_ref.greedy_decode(logits[b, :lengths[b]], V - 1) if lengths[b] else []
        assert ids[b, :lens[b]].tolist() == ref
        assert (ids[b, lens[b]:] == -1).all()
    assert lens[1] == 0 and lens[2] == 1 and lens[4] == 0
    ids0, lens0 = eng32.ctc_greedy(logits, blank=0)                 # blank_at_zero convention
    assert ids0[2, :int(lens0[2])].tolist() == []


def test_beam_kernel_vs_oracle_small(eng32):
    from oracle import ctc_ref
    rng = np.random.default_rng(9)
    for (T, V, scale, beam, cp, ctn) in [(12, 9, 2.0, 4, 1.0, 40), (30, 40, 4.0, 8, 1.0, 40), (25, 50, 3.0, 8, 0.99, 20),
                                          (40, 60, 6.0, 16, 1.0, 40), (20, 30, 1.0, 1, 1.0, 40), (18, 70, 5.0, 32, 1.0, 40)]:
        logits = (rng.standard_normal((3, T, V)) * scale).astype(np.float32)
        lengths = np.array([T, T - 3, max(T // 2, 1)], np.int32)
        ids, lens, scores = eng32.ctc_beam(logits, beam, lengths=lengths, cutoff_prob=cp, cutoff_top_n=ctn)
        ids, lens, scores = ids.cpu().numpy(), lens.cpu().numpy(), scores.cpu().numpy()
        for b in range(3):
            probs = ctc_ref.softmax(logits[b, :lengths[b]].astype(np.float32)).astype(np.float32)
            ref = ctc_ref.beam_search(probs.astype(np.float64), beam, cutoff_prob=cp, cutoff_top_n=ctn)
            n = len(ref)
            got = [ids[b, k, :lens[b, k]].tolist() for k in range(beam) if lens[b, k] >= 0]
            assert len(got) == n, (T, V, beam, len(got), n)
            # identical hypotheses in identical order unless two neighbours tie to within float rounding
            for k in range(n):
                if got[k] != ref[k][1]:
                    assert abs(ref[k][0] - scores[b, k]) < 1e-3 and sorted(map(tuple, got)) == sorted(tuple(r[1]) for r in ref)
            np.testing.assert_allclose(scores[b, :n], [r[0] for r in ref], atol=2e-3)


def test_beam_on_reference_wav(eng32, golden, ref_wav):
    logits = eng32.ctc_logits(eng32.encode(ref_wav[None]))
    ids, lens, scores = eng32.ctc_beam(logits, 16)
    ids, lens, scores = ids.cpu().numpy()[0], lens.cpu().numpy()[0], scores.cpu().numpy()[0]
    for k in range(4):
        assert ids[k, :lens[k]].tolist() == golden["wav_beam_ids"][k].tolist()
    np.testing.assert_allclose(scores[:4], golden["wav_beam_scores"], atol=2e-3)
    assert (lens >= 0).all()                                         # the reference returns 16 hypotheses too
    i1, l1, s1 = eng32.ctc_beam(logits, 1)
    assert i1[0, 0, :int(l1[0, 0])].tolist() == GOLDEN_IDS           # beam 1 == greedy invariant


# ------------------------------------------------------------------------------------------------- other geometries
def test_streaming_model_chunks(streaming_weights, golden, ref_wav):
    """StreamingConformerCTC (dmodel 256, 4 blocks, kernel 5): independent 8000-sample chunks, global CTC decoder."""
    from tensorflowasr_b200 import engine as E
    ge, re_, gc, rc = streaming_weights
    for prec, tol in ((1, 1e-3), (0, 5e-2)):
        e = E.Engine(ge, re_, gc, rc, precision=prec, chunk_samples=8000)
        full = ref_wav[:64000].reshape(1, 64000)                     # 8 whole chunks in one call (reshape path, :574-594)
        enc = e.encode(full).cpu().numpy()
        assert enc.shape == (1, 104, 256)
        np.testing.assert_allclose(enc[0], golden["stream_enc"][:104], atol=tol)
        tail = e.encode(ref_wav[None, 64000:]).cpu().numpy()          # ragged last chunk on its own (test_asr.py:120-128)
        np.testing.assert_allclose(tail[0], golden["stream_enc"][104:], atol=tol)
        allenc = np.concatenate([enc, tail], axis=1)
        ids, lens = e.ctc_greedy(e.ctc_logits(allenc))
        assert ids[0, :int(lens[0])].tolist() == golden["stream_ids"].tolist()
        e.close()


@pytest.mark.parametrize("dmodel,heads,hs,ks", [(144, 4, 36, 32), (256, 4, 64, 5), (64, 2, 32, 7)])
def test_random_weight_models(dmodel, heads, hs, ks):
    from oracle import conformer_ref as cr
    from tensorflowasr_b200 import engine as E, weights as W
    ge, re_, gc, rc = W.random_model(3, dmodel=dmodel, num_blocks=2, num_heads=heads, head_size=hs, kernel_size=ks, vocab=100)
    x = (np.random.default_rng(0).standard_normal((3, 9000)) * 0.1).astype(np.float32)
    enc_ref = cr.encoder_forward(x, re_, 2)
    lg_ref = cr.ctc_forward(enc_ref, rc, 1)
    for prec, te, tl in ((1, 1e-3, 5e-3), (0, 8e-2, 0.5)):
        e = E.Engine(ge, re_, gc, rc, precision=prec)
        enc = e.encode(x)
        np.testing.assert_allclose(enc.cpu().numpy(), enc_ref, atol=te)
        np.testing.assert_allclose(e.ctc_logits(enc).cpu().numpy(), lg_ref, atol=tl)
        e.close()


def test_abi_argument_errors(eng32, torch_mod):
    lib, h = eng32.lib, eng32._h
    assert lib.b200asr_encode(h, None, 1, 100, None, None) != 0
    assert b"bad arguments" in lib.b200asr_last_error(h)
    x = torch_mod.zeros(4, device="cuda")
    assert lib.b200asr_ctc_beam(h, x.data_ptr(), None, 1, 1, 4, 3, 99, 40, 1.0, x.data_ptr(), x.data_ptr(), x.data_ptr(), None) != 0
    assert b"beam size" in lib.b200asr_last_error(h)
    assert lib.b200asr_recognize(h, x.data_ptr(), 0, 100, x.data_ptr(), x.data_ptr(), None) == 0   # empty batch is a no-op


def test_reference_facing_asr_surface(ref_wav):
    """ASR(config).compile(dir).stt(wav) -- the call sequence of Inference/PythonInference/asr/src/asr.py / test_asr.py."""
    from oracle import ort_ref
    from tensorflowasr_b200 import asr as A
    vocab = os.path.join(ort_ref.REF_DIR, "dict", "pinyin.txt")
    if not os.path.isfile(vocab) or ort_ref.model_dir("offline") is None:
        pytest.skip("reference vocabulary / models not staged")
    cfg = {"running_config": {}, "optimizer_config": {}, "tar_config": None,
           "speech_config": {"sample_rate": 16000, "frame_ms": 25, "stride_ms": 10, "num_feature_bins": 80, "streaming": False,
                             "streaming_bucket": 0.5},
           "model_config": {"dmodel": 144, "num_blocks": 13, "num_heads": 4, "head_size": 36, "kernel_size": 32},
           "inp_config": {"vocabulary": vocab, "blank_at_zero": False, "beam_width": 1}}
    a = A.ASR(cfg)
    a.compile(ort_ref.model_dir("offline"))
    wav_path = os.path.join(os.path.dirname(__file__), "golden", "BAC009S0764W0121.wav")
    phones, text = a.stt(wav_path)
    assert phones.split(" ") == a.phone_featurizer.iextract(GOLDEN_IDS) and text == ""
    feat = a.extract_feature(ref_wav)
    assert feat.shape == (1, 106, 144) and feat.dtype == np.float32
    assert a.decode([feat]) == phones
    assert a.decode([feat[:, :50], feat[:, 50:]]) == phones               # hstack along time (asr.py:65-66)
    logits = a.engine.ctc_logits(feat).cpu().numpy()[0]
    assert a.greedy_decode(a.softmax(logits), 1331) == GOLDEN_IDS
    cfg["inp_config"]["beam_width"] = 8                                   # the YAML key the reference never reads
    b = A.ASR(cfg)
    b.compile(ort_ref.model_dir("offline"))
    assert b.stt(wav_path)[0] == phones
    assert a.recognize_batch(np.stack([ref_wav[:30000], ref_wav[:30000]]))[0] == a.recognize_batch(ref_wav[None, :30000])[0]


@pytest.mark.parametrize("pair", [False, True], ids=["chain", "chain_pair"])
@pytest.mark.parametrize("M,N1", [(8000, 576), (8000, 288), (300, 576), (1, 288), (129, 576), (8064, 288)])
def test_chained_ffn_kernels_vs_fp64(eng32, torch_mod, pair, M, N1):
    """The chained FFN / conv-tail tcgen05 kernels (hidden activations in TMEM; `pair`: hidden dimension split across a 2-CTA
    cluster with the partial sums exchanged through distributed shared memory) against torch fp64:
    C = resid + 0.5 * (swish(X W1^T + b1) W2^T + b2) with one or two fused LayerNorms.  tf32 tolerance 3e-2 on values ~20."""
    torch = torch_mod
    torch.manual_seed(M + N1)
    K1 = N2 = 144

    def ln(x, g, b, eps=1e-3):
        mu = x.mean(-1, keepdim=True)
        var = ((x - mu) ** 2).mean(-1, keepdim=True)
        return (x - mu) / torch.sqrt(var + eps) * g + b

    X = torch.randn(M, K1, device="cuda")
    W1 = torch.randn(N1, K1, device="cuda") / K1 ** 0.5
    b1 = torch.randn(N1, device="cuda") * 0.3
    W2 = torch.randn(N2, N1, device="cuda") / N1 ** 0.5
    b2 = torch.randn(N2, device="cuda") * 0.3
    g1, be1, g2, be2 = (torch.randn(N2, device="cuda") for _ in range(4))
    for epi in (6, 7):
        resid = torch.randn(M, N2, device="cuda") * 20
        hid = X.double() @ W1.double().T + b1.double()
        hid = hid * torch.sigmoid(hid)
        x = resid.double() + 0.5 * (hid @ W2.double().T + b2.double())
        if epi == 6:
            c_ref, c2_ref = x, ln(x, g1.double(), be1.double())
        else:
            c_ref = ln(x, g1.double(), be1.double())
            c2_ref = ln(c_ref, g2.double(), be2.double())
        C, C2 = eng32.debug_chain(X, W1, b1, W2, b2, resid, 0.5, epi, (g1, be1), (g2, be2) if epi == 7 else None, pair=pair)
        torch.cuda.synchronize()
        assert not torch.isnan(C2).any()
        assert (C.double() - c_ref).abs().max().item() < 3e-2
        assert (C2.double() - c2_ref).abs().max().item() < 3e-2


def test_host_pipeline_matches_synchronous_call(eng, ref_wav, torch_mod):
    """b200asr_recognize_host_submit / _collect (two slots, H2D of one batch under the compute of the other) returns exactly
    what the synchronous host call returns, for alternating inputs and re-used slots."""
    torch = torch_mod
    L = 24000
    rng = np.random.default_rng(11)
    batches = []
    for i in range(5):
        x = np.clip(rng.standard_normal((3, L)).astype(np.float32) * 0.1, -1, 1)
        x[i % 3] = np.tile(ref_wav, 2)[i * 1000:i * 1000 + L]
        batches.append(torch.from_numpy(x).pin_memory())
    want = []
    for xb in batches:
        ids, lens = eng.recognize_host(xb)
        want.append((ids.clone(), lens.clone()))
    Tp = want[0][0].shape[1]
    hid = [torch.empty((3, Tp), dtype=torch.int32).pin_memory() for _ in range(2)]
    hlen = [torch.empty((3,), dtype=torch.int32).pin_memory() for _ in range(2)]
    got = [None] * len(batches)
    for i, xb in enumerate(batches):
        sl = i & 1
        if i >= 2:
            eng.recognize_host_collect(sl)
            got[i - 2] = (hid[sl].clone(), hlen[sl].clone())
        eng.recognize_host_submit(sl, xb, hid[sl], hlen[sl])
    for i in range(len(batches) - 2, len(batches)):
        eng.recognize_host_collect(i & 1)
        got[i] = (hid[i & 1].clone(), hlen[i & 1].clone())
    for (wi, wl), (gi, gl) in zip(want, got):
        assert torch.equal(wl, gl)
        for b in range(3):
            n = int(wl[b])
            assert torch.equal(wi[b, :n], gi[b, :n])
    with pytest.raises(RuntimeError):
        eng.recognize_host_collect(0)          # nothing in flight any more


def test_fused_argmax_head_equals_logits_path(eng, ref_wav, torch_mod):
    """recognize() (tf32: CTC head fused with the per-frame argmax, no logits materialised) returns exactly the ids obtained by
    collapsing the argmax of the logits the same engine writes through b200asr_ctc_logits -- the reference's greedy rule
    (first maximum wins, ctc_greedy_decoder.h:11-18) on identical arithmetic."""
    from oracle import ctc_ref
    L = 40000
    rng = np.random.default_rng(5)
    x = np.clip(rng.standard_normal((6, L)).astype(np.float32) * 0.1, -1, 1)
    x[1] = np.tile(ref_wav, 2)[:L]
    x[4] = np.tile(ref_wav, 2)[3000:3000 + L]
    ids, lens = eng.recognize(x)
    logits = eng.ctc_logits(eng.encode(x)).cpu().numpy()
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    for b in range(x.shape[0]):
        want = ctc_ref.greedy_decode(logits[b], logits.shape[-1] - 1)
        assert ids[b, :lens[b]].tolist() == want
    assert lens[1] > 5 and lens[4] > 5


def test_config5_full_size_beam16_properties(eng, ref_wav, torch_mod):
    """BASELINE config 5 shape (ConformerCTC(S) + prefix beam 16, batch 128 x 5 s): size-independent properties -- hypotheses
    sorted by score, beam 1 == greedy on every row, tiled-speech rows all return the same 16 hypotheses, best beam hypothesis ==
    greedy ids where the greedy path is unambiguous (speech rows)."""
    L, B = 80000, 128
    speech = np.tile(ref_wav, 2)[:L]
    rng = np.random.default_rng(1237)
    x = np.clip(rng.standard_normal((B, L)).astype(np.float32) * 0.1, -1, 1)
    x[::8] = speech
    xs = torch_mod.from_numpy(x).cuda()
    gids, glens = eng.recognize(xs)
    logits = eng.ctc_logits(eng.encode(xs))
    ids, lens, scores = eng.ctc_beam(logits, 16)
    i1, l1, _ = eng.ctc_beam(logits, 1)
    gids, glens, ids, lens, scores = (t.cpu().numpy() for t in (gids, glens, ids, lens, scores))
    i1, l1 = i1.cpu().numpy(), l1.cpu().numpy()
    assert np.isfinite(scores[lens >= 0]).all()
    for b in range(B):
        n = int((lens[b] >= 0).sum())
        assert n >= 1 and (np.diff(scores[b, :n]) <= 1e-6).all()                       # descending scores
        assert i1[b, 0, :l1[b, 0]].tolist() == gids[b, :glens[b]].tolist()             # beam 1 == greedy
    for b in range(0, B, 8):
        assert (lens[b] == lens[0]).all()                                              # identical rows, identical beams
        for k in range(16):
            if lens[0, k] >= 0:
                assert (ids[b, k, :lens[b, k]] == ids[0, k, :lens[0, k]]).all()
        assert ids[b, 0, :lens[b, 0]].tolist() == gids[b, :glens[b]].tolist()
    assert glens[0] >= 13                                                               # >= one pass of the 13-token utterance


def test_config3_full_size_streaming_properties(streaming_weights, ref_wav, torch_mod):
    """BASELINE config 3 shape (StreamingConformerCTC, batch 64 x 30 s = 3840 independent 8000-sample chunks, global CTC
    decoder over 780 frames): finite, identical rows decode identically, and a row equals the same utterance run alone."""
    from tensorflowasr_b200 import engine as E
    ge, re_, gc, rc = streaming_weights
    e = E.Engine(ge, re_, gc, rc, precision=0, chunk_samples=8000)
    L, B = 480000, 64
    speech = np.tile(ref_wav, 8)[:L]
    rng = np.random.default_rng(1235)
    x = np.clip(rng.standard_normal((B, L)).astype(np.float32) * 0.1, -1, 1)
    x[::16] = speech
    xs = torch_mod.from_numpy(x).cuda()
    ids, lens = e.recognize(xs)
    enc = e.encode(xs)
    assert tuple(enc.shape) == (64, 780, 256) and torch_mod.isfinite(enc).all()
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    one_ids, one_len = e.recognize(speech[None])
    for r in range(0, B, 16):
        assert lens[r] == int(one_len[0]) and (ids[r, :lens[r]] == one_ids[0, :lens[r]].cpu().numpy()).all()
    assert lens[0] >= 60
    e.close()


def test_translator_vs_reference_golden(offline_weights, golden):
    """b200asr_translate against the reference's translator.onnx on the reference wav: per-position character argmax
    identical in both precision modes, logits within 2e-3 (fp32 mode) / the tf32 outlier bound 0.25 (measured 0.11)."""
    import os
    from oracle import ort_ref, conformer_ref as cr
    from tensorflowasr_b200 import engine as E, weights as W
    md = ort_ref.model_dir("offline")
    if md is None or not os.path.isfile(os.path.join(md, "translator.onnx")):
        pytest.skip("translator.onnx not staged")
    ge, re_, gc, rc = offline_weights
    gt, rt = W.import_translator(os.path.join(md, "translator.onnx"))
    ref = cr.translator_forward(golden["wav_tr_in"][None], golden["wav_enc"][None], rt, gt.num_blocks)
    for prec, tol in ((1, 2e-3), (0, 0.25)):        # tf32: the outlier bound of tests/conftest.py (logits up to ~40); argmax must be identical
        e = E.Engine(ge, re_, gc, rc, precision=prec, tr_geo=gt, tr_raw=rt)
        out = e.translate(golden["wav_tr_in"][None], golden["wav_enc"][None]).cpu().numpy()
        assert out.shape == (1, 23, 9160)
        err = float(np.abs(out - ref).max())
        print(f"translator precision {prec}: max |logits - oracle| = {err:.3e}")
        assert (out[0].argmax(-1) == golden["wav_tr_argmax"]).all()
        np.testing.assert_allclose(out[0, :4], golden["wav_tr_logits_rows"], atol=tol)
        assert err <= tol
        two = e.translate(np.stack([golden["wav_tr_in"], golden["wav_tr_in"][::-1].copy()]), np.stack([golden["wav_enc"]] * 2)).cpu().numpy()
        np.testing.assert_allclose(two[0], out[0], atol=1e-4 if prec else 2e-2)            # batch rows are independent
        e.close()


def test_asr_surface_returns_text(ref_wav):
    """decode() / stt() return what the reference returns: characters from the translator (asr.py:62-94; test_asr.py:186-218)."""
    from oracle import ort_ref
    from tensorflowasr_b200 import asr as A
    d = ort_ref.model_dir("offline")
    vocab = os.path.join(ort_ref.REF_DIR, "dict", "pinyin.txt")
    lm = os.path.join(ort_ref.REF_DIR, "dict", "lm_tokens.txt")
    if d is None or not all(os.path.isfile(p) for p in (vocab, lm, os.path.join(d, "translator.onnx"))):
        pytest.skip("translator / vocabularies not staged")
    cfg = {"running_config": {}, "optimizer_config": {},
           "speech_config": {"sample_rate": 16000, "frame_ms": 25, "stride_ms": 10, "num_feature_bins": 80, "streaming": False,
                             "streaming_bucket": 0.5},
           "model_config": {"dmodel": 144, "num_blocks": 13, "num_heads": 4, "head_size": 36, "kernel_size": 32},
           "inp_config": {"vocabulary": vocab, "blank_at_zero": False, "beam_width": 1},
           "tar_config": {"vocabulary": lm, "blank_at_zero": False, "beam_width": 1}}
    a = A.ASR(cfg)
    a.compile(d)
    assert a.has_translator()
    feat = a.extract_feature(ref_wav)
    assert a.decode([feat]) == "甚至出现交易几乎停制的情况"
    wav_path = os.path.join(os.path.dirname(__file__), "golden", "BAC009S0764W0121.wav")
    phones, text = a.stt(wav_path)
    assert phones.split(" ") == a.phone_featurizer.iextract(GOLDEN_IDS)
    assert text.startswith("甚至出现交易几乎停制的情况")


def test_beam_probability_input_is_bit_exact(eng32):
    """b200asr_ctc_beam_probs against the reference's own C++ decoder (oracle/_ref/libctcdec_ref.so) on IDENTICAL probabilities: the same
    hypotheses in the same order and float scores equal to within 2 ulp (> 95 % bit-identical; only a tie of score and last token, which
    the reference itself leaves to an unstable sort, may be ordered either way) -- small beams on peaky / flat distributions included
    (the case where a pruned prefix with live children is revived, path_trie.cpp:37-51)."""
    from oracle import ctc_ref, ctcdec_ref
    if not ctcdec_ref.available():
        pytest.skip("oracle/_ref/libctcdec_ref.so not staged")
    rng = np.random.default_rng(21)
    cases = [(40, 12, 3.0, 2), (40, 12, 3.0, 3), (60, 8, 1.0, 2), (60, 8, 0.3, 4), (50, 30, 5.0, 4), (80, 1332, 6.0, 16), (125, 1332, 2.0, 16),
             (30, 6, 0.1, 3), (64, 20, 2.0, 8), (33, 50, 4.0, 32)]
    n_hyp = n_ties = n_exact = 0
    for (T, V, scale, beam) in cases:
        logits = (rng.standard_normal((4, T, V)) * scale).astype(np.float32)
        logits[1, :, V - 1] += 2.0                                             # blank-heavy
        logits[2, ::3, 2] += 4.0                                               # frequent repeats of one token
        probs = np.stack([ctc_ref.softmax(l) for l in logits]).astype(np.float32)
        ids, lens, scores = eng32.ctc_beam(probs, beam, probs=True)
        ids, lens, scores = ids.cpu().numpy(), lens.cpu().numpy(), scores.cpu().numpy()
        for b in range(4):
            ref = ctcdec_ref.beam_search(probs[b].astype(np.float64), beam)
            got = [ids[b, k, :lens[b, k]].tolist() for k in range(beam) if lens[b, k] >= 0]
            if got != [r[1] for r in ref]:          # diagnostics: where, with which scores, and is it a reordering or a different set
                refs = [r[1] for r in ref]
                for k in range(max(len(got), len(refs))):
                    gk = got[k] if k < len(got) else None
                    rk = refs[k] if k < len(refs) else None
                    if gk != rk:
                        print(f"beam mismatch case {(T, V, scale, beam, b)} at rank {k}: device score {scores[b, k]!r} (this hypothesis is reference rank "
                              f"{refs.index(gk) if gk in refs else None}), reference score {np.float32(ref[k][0])!r} (reference hypothesis is device rank "
                              f"{got.index(rk) if rk in got else None}); lengths {len(gk) if gk else None} / {len(rk) if rk else None}")
            rs = np.asarray([r[0] for r in ref], dtype=np.float32)
            gs = scores[b, :len(ref)]
            ulp = np.abs(gs.view(np.int32).astype(np.int64) - rs.view(np.int32).astype(np.int64))
            assert ulp.max() <= 2, (T, V, scale, beam, b, ulp.tolist())     # float scores equal to the last bit or two (glibc expf / logf
            n_exact += int((ulp == 0).sum())                                # are not always correctly rounded; the device evaluates in double)
            # order: identical, except inside a group of hypotheses whose reference scores are equal (to those 2 ulp) and whose last tokens
            # agree, which the reference's prefix_compare leaves to its unstable std::sort (decoder_utils.cpp:137-147; even the pinned CPU
            # restatement orders such a pair differently from the C++ build)
            k = 0
            while k < len(ref):
                e = k + 1
                while e < len(ref) and abs(int(rs[e:e + 1].view(np.int32)[0]) - int(rs[k:k + 1].view(np.int32)[0])) <= 2:
                    e += 1
                if e - k == 1:
                    assert got[k] == ref[k][1], (T, V, scale, beam, b, k)
                else:
                    assert sorted(map(tuple, got[k:e])) == sorted(tuple(r[1]) for r in ref[k:e]), (T, V, scale, beam, b, k)
                    n_ties += e - k
                k = e
            n_hyp += len(ref)
    print(f"beam, probability input: {n_hyp} hypotheses, {n_exact} scores bit-identical, the rest within 2 ulp; {n_ties} inside tie groups")
    assert n_hyp > 200 and n_ties < n_hyp // 10 and n_exact > 0.95 * n_hyp
