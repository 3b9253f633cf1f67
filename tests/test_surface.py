"""CPU tests of the reference-facing Python surface (config / vocabulary / wav loading) and of the multi-process
utterance sharding logic (gloo, world_size 2)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from tensorflowasr_b200 import asr as A, sharding

REF_CFG = "/root/reference/asr/configs"


def _vocab_file(tmp_path, tokens):
    p = tmp_path / "vocab.txt"
    p.write_text("# comment\n" + "\n".join(tokens) + "\n", encoding="utf-8")
    return str(p)


def test_text_featurizer_blank_last(tmp_path):
    tf = A.TextFeaturizer({"vocabulary": _vocab_file(tmp_path, ["<S>", "</S>", "[SPACE]", "a1", "b2"]), "blank_at_zero": False})
    assert tf.num_classes == 6 and tf.blank == 5
    assert tf.startid() == 0 and tf.endid() == 1 and tf.token_to_index[" "] == 2
    assert tf.iextract([3, 4]) == ["a1", "b2"] and tf.extract(["b2"]) == [4]
    tz = A.TextFeaturizer({"vocabulary": _vocab_file(tmp_path, ["x", "y"]), "blank_at_zero": True})
    assert tz.blank == 0 and tz.num_classes == 3 and tz.token_to_index["x"] == 1


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference configs not present")
def test_reference_config_files_load_unchanged():
    cfg = A.UserConfig(os.path.join(REF_CFG, "am_data.yml"), os.path.join(REF_CFG, "conformerS.yml"))
    assert cfg["speech_config"]["stride_ms"] == 10 and cfg["model_config"]["dmodel"] == 144
    assert cfg["inp_config"]["blank_at_zero"] is False and cfg["inp_config"]["beam_width"] == 1
    assert cfg["no_such_key"] is None                                   # UserDict.__missing__ (utils/user_config.py:24-25)
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        tf = A.TextFeaturizer(dict(cfg["inp_config"]))
    finally:
        os.chdir(cwd)
    assert tf.num_classes == 1332 and tf.blank == 1331
    assert " ".join(tf.iextract([669, 82, 103])) == "shen4 zhi4 chu1" or len(tf.iextract([669, 82, 103])) == 3


def test_load_wav_matches_int16_scaling(ref_wav):
    sf = A.SpeechFeaturizer({"sample_rate": 16000, "frame_ms": 25, "stride_ms": 10, "num_feature_bins": 80})
    x = sf.load_wav(os.path.join(GOLDEN, "BAC009S0764W0121.wav"))
    assert x.dtype == np.float32 and x.shape == (67263,)
    np.testing.assert_array_equal(x, ref_wav)
    padded = sf.pad_signal([x[:10], x[:4]], 6)
    assert padded.shape == (2, 6) and (padded[1, 4:] == 0).all() and (padded[0] == x[:6]).all()


def test_shard_range_covers_batch():
    for B in (0, 1, 7, 32, 129):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeEngine:                      # stands in for the GPU engine: "ids" encode (global utterance index, length)
        def recognize(self, wavs):
            n = len(wavs)
            T = 5 + rank                   # ragged T across ranks
            ids = torch.full((n, T), -1, dtype=torch.int32)
            lens = torch.zeros((n,), dtype=torch.int32)
            for i in range(n):
                k = int(wavs[i][0]) % 4 + 1
                ids[i, :k] = int(wavs[i][0])
                lens[i] = k
            return ids, lens

    B = 7
    wavs = np.arange(B, dtype=np.float32).reshape(B, 1).repeat(3, axis=1)
    ids, lens = sharding.recognize_sharded(FakeEngine(), wavs)
    ok = ids.shape == (B, 6) and all(int(lens[i]) == i % 4 + 1 and (ids[i, :int(lens[i])] == i).all() and
                                     (ids[i, int(lens[i]):] == -1).all() for i in range(B))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_recognize_gloo_world2():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 500
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] is True and out[1] is True


def _xch_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, T = 3, 5
    xch = sharding.IdsExchange(B, T, torch.device("cpu"), slots=2)
    ok = True
    for step in range(5):                       # more steps than slots: slots are re-acquired after their collective
        sl = xch.acquire()
        ids, lens = xch.buffers(sl)
        ids.fill_(-1)
        for b in range(B):
            n = (rank + b + step) % T
            ids[b, :n] = 100 * rank + 10 * step + b
            lens[b] = n
        xch.gather(sl)
        all_ids, all_lens = xch.result(sl)
        ok = ok and tuple(all_ids.shape) == (world * B, T) and tuple(all_lens.shape) == (world * B,)
        for r in range(world):
            for b in range(B):
                n = (r + b + step) % T
                ok = ok and int(all_lens[r * B + b]) == n and bool((all_ids[r * B + b, :n] == 100 * r + 10 * step + b).all()) and \
                    bool((all_ids[r * B + b, n:] == -1).all())
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_ids_exchange_single_collective_gloo_world2():
    """sharding.IdsExchange: ids and lengths travel in ONE all_gather_into_tensor of a flat buffer the decoder writes in place."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() + 137) % 500
    mp.spawn(_xch_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] is True and out[1] is True
