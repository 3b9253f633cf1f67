"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference's CTC decoders.

greedy_decode: Inference/PythonInference/asr/src/asr.py:41-61 == test_asr.py:167-185 ==
  Inference/CppInference/onnx/src/core/ctc_greedy_decoder.h:4-43 (argmax per frame, merge repeats, drop blank).
beam_search:   externals/ctc_decoders.zip: ctc_beam_search_decoder.cpp:18-187, decoder_utils.cpp:7-38,137-147,
  decoder_utils.h:42-49, path_trie.cpp:37-158 with ext_scorer == nullptr, restated over integer prefixes.
  Pinned against the reference's own C++ (oracle/_ref/libctcdec_ref.so, see ctcdec_ref.py) in tests/test_oracle.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NUM_FLT_INF = float(np.finfo(np.float32).max)
NUM_FLT_MIN = float(np.finfo(np.float32).tiny)      # std::numeric_limits<float>::min()


def softmax(logits: np.ndarray) -> np.ndarray:
    """asr.py:27-33 (numpy softmax over the last axis)."""
    x = np.asarray(logits, dtype=np.float64)
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def greedy_decode(y: np.ndarray, blank: int) -> List[int]:
    raw = np.argmax(y, axis=1)            # first maximum wins (numpy and ctc_greedy_decoder.h:11-18 agree)
    out: List[int] = []
    prev = None
    for s in raw.tolist():
        if s != prev:
            if s != blank:
                out.append(int(s))
            prev = s
    return out


def _f32(x: float) -> float:
    return float(np.float32(x))


def log_sum_exp(x: float, y: float) -> float:
    """decoder_utils.h:42-49 on float (T = float at every call site of the search)."""
    if x <= -NUM_FLT_INF:
        return y
    if y <= -NUM_FLT_INF:
        return x
    m = max(x, y)
    return _f32(_f32(math.log(_f32(math.exp(_f32(x - m))) + _f32(math.exp(_f32(y - m))))) + m)


def pruned_log_probs(prob_step: np.ndarray, cutoff_prob: float, cutoff_top_n: int) -> List[Tuple[int, float]]:
    """decoder_utils.cpp:7-38, faithfully: when cutoff_prob == 1.0 the vocabulary is only *sorted* (cutoff_len stays
    = V, so cutoff_top_n has no effect); the cumulative/top-n cut applies only for cutoff_prob < 1.0.
    (std::sort's order among equal probabilities is unspecified; ties are broken by index here.)"""
    V = len(prob_step)
    order = list(range(V))
    cutoff_len = V
    if cutoff_prob < 1.0 or cutoff_top_n < cutoff_len:
        order = sorted(range(V), key=lambda i: (-float(prob_step[i]), i))
        if cutoff_prob < 1.0:
            cum = 0.0
            cutoff_len = 0
            for i in range(V):
                cum += float(prob_step[order[i]])
                cutoff_len += 1
                if cum >= cutoff_prob or cutoff_len >= cutoff_top_n:
                    break
        order = order[:cutoff_len]
    return [(i, _f32(math.log(float(prob_step[i]) + NUM_FLT_MIN))) for i in order]


class _Prefix:
    __slots__ = ("tokens", "b_prev", "nb_prev", "b_cur", "nb_cur", "score")

    def __init__(self, tokens):
        self.tokens = tokens
        self.b_prev = self.nb_prev = self.b_cur = self.nb_cur = self.score = -NUM_FLT_INF


def beam_search(probs: np.ndarray, beam_size: int, blank: Optional[int] = None, cutoff_prob: float = 1.0,
                cutoff_top_n: int = 40) -> List[Tuple[float, List[int]]]:
    """ctc_beam_search_decoder (ctc_beam_search_decoder.cpp:18-187) without scorer; probs [T, V] post-softmax with the
    blank as the last class (the reference passes vocabulary of size V-1, blank_id = V-1, :30).  Returns up to
    beam_size (log_prob, ids) sorted by score descending."""
    T, V = probs.shape
    blank = V - 1 if blank is None else blank
    root = _Prefix(())
    root.score = root.b_prev = 0.0
    prefixes: Dict[tuple, _Prefix] = {(): root}
    live = [root]
    for t in range(T):
        step = pruned_log_probs(probs[t], cutoff_prob, cutoff_top_n)
        # (no scorer => no min_cutoff pruning: full_beam stays false, :64-72)
        created: List[_Prefix] = []
        for c, log_prob_c in step:
            for p in live[:beam_size]:
                if c == blank:
                    p.b_cur = log_sum_exp(p.b_cur, _f32(log_prob_c + p.score))
                    continue
                if p.tokens and c == p.tokens[-1]:
                    p.nb_cur = log_sum_exp(p.nb_cur, _f32(log_prob_c + p.nb_prev))
                key = p.tokens + (c,)
                q = prefixes.get(key)
                if q is None:
                    q = _Prefix(key)
                    prefixes[key] = q
                    created.append(q)
                if p.tokens and c == p.tokens[-1] and p.b_prev > -NUM_FLT_INF:
                    log_p = _f32(log_prob_c + p.b_prev)
                elif not (p.tokens and c == p.tokens[-1]):
                    log_p = _f32(log_prob_c + p.score)
                else:
                    log_p = -NUM_FLT_INF
                q.nb_cur = log_sum_exp(q.nb_cur, log_p)
        # iterate_to_vec (path_trie.cpp:112-126): roll cur -> prev for every node that exists
        live = []
        for p in prefixes.values():
            p.b_prev, p.nb_prev = p.b_cur, p.nb_cur
            p.b_cur = p.nb_cur = -NUM_FLT_INF
            p.score = log_sum_exp(p.b_prev, p.nb_prev)
            live.append(p)
        if len(live) >= beam_size:
            live.sort(key=lambda p: (-p.score, p.tokens))            # nth_element + prefix_compare
            for p in live[beam_size:]:
                _remove(prefixes, p)
            live = live[:beam_size]
    live = [p for p in live if True]
    live.sort(key=lambda p: (-p.score, p.tokens[-1] if p.tokens else -1))
    return [(p.score, list(p.tokens)) for p in live[:beam_size]]


def _remove(prefixes: Dict[tuple, _Prefix], p: _Prefix):
    """PathTrie::remove (path_trie.cpp:128-146): a pruned node disappears only if it has no children; otherwise it
    stays in the trie (exists_ = false) -- it can never be extended again because it is not in the live beam, but its
    descendants keep their own state."""
    prefixes.pop(p.tokens, None)
