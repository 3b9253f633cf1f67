// CTC prefix beam search on the device, no external scorer.
//
// Reference algorithm: externals/ctc_decoders.zip ctc_beam_search_decoder.cpp:18-187 (+ get_pruned_log_probs
// decoder_utils.cpp:7-38, log_sum_exp decoder_utils.h:42-49, PathTrie path_trie.cpp:37-146, prefix_compare
// decoder_utils.cpp:137-147), which walks a pointer trie on the host with one thread per utterance.  Here:
//   * beam_prep_kernel   (one warp per frame, all frames in parallel): fp32 softmax statistics and the sorted
//     top-N classes of the frame with their float log-probabilities log(double(p) + FLT_MIN);
//   * beam_search_kernel (one CTA per utterance, sequential in time): the live prefixes sit in shared memory; a
//     prefix is identified by (parent node id, last token) instead of a trie pointer; each step scores
//     beam x (beam+1) extensions + beam survivors, sorts them with a bitonic network on 64-bit keys
//     (score desc, last token asc = prefix_compare) and keeps the best `beam`; token strings are rebuilt at the
//     end from per-step back-pointers.
// Only the top beam+1 non-blank classes of a frame can produce a surviving *new* prefix (any other extension is
// dominated by beam better ones from the same parent), so the full-vocabulary double loop of the reference is
// not needed; blank / repeated-token updates use the exact class log-probability whatever its rank.
// When cutoff_prob == 1.0 the reference does not prune the vocabulary at all (cutoff_top_n is ignored, see
// decoder_utils.cpp:17-31); cutoff_prob < 1.0 restricts every update to the first cutoff_len sorted classes.
#include "kernels.cuh"

#include <float.h>

namespace b200asr {

namespace {

constexpr int kMaxBeam = 32;
constexpr int kMaxTop = 72;   // >= 2 * kMaxBeam + 2 sorted classes per frame (see beam_n_store)
constexpr float kNegInf = -FLT_MAX;  // NUM_FLT_INF of the reference is FLT_MAX

struct FrameTop {   // per frame, produced by beam_prep_kernel
  float row_max, row_sum;
  int n_valid;      // classes allowed this step (V when cutoff_prob >= 1)
  int n_top;        // entries stored below
};

// is_prob: the input rows already hold probabilities (the reference decoder's own input, probs_seq); else logits, soft-maxed here in fp32
__device__ __forceinline__ float class_logprob(float x, float row_max, float row_sum, int is_prob) {
  const float p = is_prob ? x : expf(x - row_max) / row_sum;
  return (float)log((double)p + (double)FLT_MIN);              // decoder_utils.cpp:33-36: log of a double, stored as float
}

// float log_sum_exp (decoder_utils.h:42-49).  The reference instantiates it with T = float: std::exp / std::log are glibc's expf /
// logf there, which return the correctly rounded float in all but vanishingly rare cases; CUDA's expf / logf do not (2 / 1 ulp).  The
// same values are produced here by evaluating in double and rounding once, so scores agree with the reference to the last bit or two (> 95 % of them bit for bit) and
// hypotheses can only swap on exact ties -- which prefix_compare resolves the same way (decoder_utils.cpp:137-147).
__device__ __forceinline__ float lse(float x, float y) {
  if (x <= kNegInf) return y;
  if (y <= kNegInf) return x;
  const float m = fmaxf(x, y);
  const float ex = (float)exp((double)(x - m)), ey = (float)exp((double)(y - m));
  return (float)log((double)(ex + ey)) + m;
}

// prefix identity: a hash of the path (root constant, then mixed with every token).  The reference identifies a prefix by its trie
// node, and a node that was pruned while it still had live children is REVIVED -- same node -- when its parent extends into it again
// (PathTrie::get_path_trie / remove, path_trie.cpp:37-51,134-152); a path hash gives a re-created prefix its old identity too, so its
// live children keep recognising it as their parent.
__device__ __forceinline__ unsigned long long mix_id(unsigned long long parent, int token) {
  unsigned long long z = parent + 0x9E3779B97F4A7C15ull * (unsigned long long)(token + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// ---------------------------------------------------------------------------------------------- prep
__global__ void __launch_bounds__(256) beam_prep_kernel(const float* __restrict__ logits, int rows, int V, int n_store,
                                                        float cutoff_prob, int cutoff_top_n, FrameTop* __restrict__ meta,
                                                        int* __restrict__ top_idx, float* __restrict__ top_lp, int is_prob) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = logits + (size_t)row * V;
  float mx = -INFINITY;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, x[v]);
  mx = warp_max(mx);
  float sm = 0.f;
  if (!is_prob) {
    for (int v = lane; v < V; v += 32) sm += expf(x[v] - mx);
    sm = warp_sum(sm);
  }
  // N rounds of "best class strictly after the previous pick" in (value desc, index asc) order
  float pv = INFINITY;
  int pi = -1;
  float cum = 0.f;
  int n_valid = V;
  bool cut_done = !(cutoff_prob < 1.0f);
  int stored = 0;
  for (int r = 0; r < n_store; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
      const float xv = x[v];
      const bool after = (xv < pv) || (xv == pv && v > pi);
      if (after && (xv > bv || (xv == bv && v < bi))) {
        bv = xv;
        bi = v;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (bi == 0x7fffffff) break;
    pv = bv;
    pi = bi;
    const float p = is_prob ? bv : expf(bv - mx) / sm;
    if (lane == 0) {
      top_idx[(size_t)row * n_store + r] = bi;
      top_lp[(size_t)row * n_store + r] = (float)log((double)p + (double)FLT_MIN);
    }
    stored = r + 1;
    if (!cut_done) {
      cum += p;
      if (cum >= cutoff_prob || (r + 1) >= cutoff_top_n) {
        n_valid = r + 1;
        cut_done = true;
      }
    }
  }
  if (!cut_done) n_valid = stored;  // cutoff_prob < 1 but the stored list ran out first (cutoff_top_n > n_store cannot happen)
  if (lane == 0) {
    FrameTop m;
    m.row_max = mx;
    m.row_sum = sm;
    m.n_valid = n_valid;
    m.n_top = stored;
    meta[row] = m;
  }
}

// ---------------------------------------------------------------------------------------------- search
struct Entry {
  float b_prev, nb_prev, score, b_cur, nb_cur;
  int last, len;
  unsigned long long id, parent_id;
};

__device__ __forceinline__ unsigned int order_key(float f) {  // larger float -> smaller key (ascending sort = score desc)
  unsigned int u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~u;
}

template <int NSORT>
__global__ void __launch_bounds__(256) beam_search_kernel(const float* __restrict__ logits, const int* __restrict__ lengths,
                                                          const FrameTop* __restrict__ meta, const int* __restrict__ top_idx,
                                                          const float* __restrict__ top_lp, int T, int V, int blank, int beam,
                                                          int n_store, int2* __restrict__ backptr /*[B,T,beam]*/,
                                                          int* __restrict__ ids, int* __restrict__ out_len,
                                                          float* __restrict__ scores, int is_prob) {
  __shared__ Entry cur[kMaxBeam];
  __shared__ Entry nxt[kMaxBeam];
  __shared__ unsigned long long keys[NSORT];
  __shared__ float cand_logp[kMaxBeam * (kMaxBeam + 1)];
  __shared__ int s_top_idx[kMaxTop];
  __shared__ float s_top_lp[kMaxTop];
  __shared__ int cand_cls[kMaxBeam * (kMaxBeam + 1)];   // class of candidate slot i * CW + j (-1 = empty)
  __shared__ int s_nbeam;

  const int b = blockIdx.x, tid = threadIdx.x;
  const int len = lengths ? min(lengths[b], T) : T;
  const int CW = beam + 1;  // children considered per prefix
  if (tid == 0) {
    Entry e;
    e.b_prev = 0.f; e.nb_prev = kNegInf; e.score = 0.f; e.b_cur = kNegInf; e.nb_cur = kNegInf;
    e.last = -1; e.id = 0x5851F42D4C957F2Dull; e.parent_id = 0ull; e.len = 0;
    cur[0] = e;
    s_nbeam = 1;
  }
  __syncthreads();

  for (int t = 0; t < len; ++t) {
    const size_t row = (size_t)b * T + t;
    const FrameTop fm = meta[row];
    const float* x = logits + row * V;
    const int nbeam = s_nbeam;
    if (tid < fm.n_top) {
      s_top_idx[tid] = top_idx[row * n_store + tid];
      s_top_lp[tid] = top_lp[row * n_store + tid];
    }
    __syncthreads();
    // 1. survivors: blank and repeated-token transitions (ctc_beam_search_decoder.cpp:88-99), plus -- when the parent of
    //    a live prefix is itself live -- the parent's extension INTO this prefix (:100-131 reaches an existing trie node),
    //    whatever the rank of the token in this frame.
    if (tid < nbeam) {
      Entry& p = cur[tid];
      auto allowed = [&](int c) -> bool {
        if (fm.n_valid >= V) return true;
        for (int r = 0; r < fm.n_valid; ++r)
          if (s_top_idx[r] == c) return true;
        return false;
      };
      p.b_cur = allowed(blank) ? class_logprob(x[blank], fm.row_max, fm.row_sum, is_prob) + p.score : kNegInf;
      float nb = kNegInf;
      if (p.last >= 0 && allowed(p.last)) {
        const float lp = class_logprob(x[p.last], fm.row_max, fm.row_sum, is_prob);
        nb = lp + p.nb_prev;
        for (int e = 0; e < nbeam; ++e) {
          const Entry& par = cur[e];
          if (par.id == p.parent_id) {
            float log_p = kNegInf;
            if (par.last == p.last) {
              if (par.b_prev > kNegInf) log_p = lp + par.b_prev;
            } else {
              log_p = lp + par.score;
            }
            nb = lse(nb, log_p);
          }
        }
      }
      p.nb_cur = nb;
    }
    // 2. extensions that create NEW prefixes (:100-131).  candidate slot = i * CW + j.  Per parent, the classes are walked in
    //    probability order and the first CW that really create a new prefix are taken: blank, classes that lead into a LIVE child of this
    //    parent (folded into that child above) and a repeat of the parent's last token without blank mass are stepped over -- so a
    //    parent with k live children still offers its best CW new extensions (the sorted list holds 2 beam + 2 classes for that).
    const int ncand_child = nbeam * CW;
    for (int k = tid; k < ncand_child; k += blockDim.x) {
      cand_logp[k] = kNegInf;
      cand_cls[k] = -1;
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      const int lim = min(fm.n_top, fm.n_valid);
      for (int i = warp; i < nbeam; i += (int)(blockDim.x >> 5)) {
        const Entry& p = cur[i];
        int taken = 0;
        for (int r0 = 0; r0 < lim && taken < CW; r0 += 32) {
          const int r = r0 + lane;
          bool ok = false;
          int c = -1;
          float log_p = kNegInf;
          if (r < lim) {
            c = s_top_idx[r];
            const float lp = s_top_lp[r];
            ok = (c != blank);
            if (ok) {
              if (c == p.last) {
                if (p.b_prev > kNegInf) log_p = lp + p.b_prev;
                else ok = false;
              } else {
                log_p = lp + p.score;
              }
            }
            if (ok)
              for (int e = 0; e < nbeam; ++e)
                if (cur[e].parent_id == p.id && cur[e].last == c) ok = false;     // a live child: already updated in step 1
          }
          const unsigned int m = __ballot_sync(0xffffffffu, ok);
          const int j = taken + __popc(m & ((1u << lane) - 1u));
          if (ok && j < CW) {
            cand_logp[i * CW + j] = log_p;
            cand_cls[i * CW + j] = c;
          }
          taken += __popc(m);
        }
      }
    }
    __syncthreads();
    // 3. keys: survivors first (index < nbeam), then children
    const int total = nbeam + ncand_child;
    for (int k = tid; k < NSORT; k += blockDim.x) {
      unsigned long long key = ~0ull;
      if (k < nbeam) {
        Entry& p = cur[k];
        p.score = lse(p.b_cur, p.nb_cur);  // iterate_to_vec (path_trie.cpp:112-121)
        if (p.score > kNegInf)
          key = ((unsigned long long)order_key(p.score) << 32) | ((unsigned long long)((p.last + 1) & 0xffff) << 16) |
                (unsigned long long)k;
      } else if (k < total) {
        const int ck = k - nbeam;
        const float s = cand_logp[ck];
        if (s > kNegInf) {
          const int c = cand_cls[ck];
          key = ((unsigned long long)order_key(s) << 32) | ((unsigned long long)((c + 1) & 0xffff) << 16) |
                (unsigned long long)k;
        }
      }
      keys[k] = key;
    }
    __syncthreads();
    // 4. bitonic sort ascending
    for (int size = 2; size <= NSORT; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int k = tid; k < NSORT / 2; k += blockDim.x) {
          const int lo = 2 * k - (k & (stride - 1));
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const unsigned long long a = keys[lo], c = keys[hi];
          if ((a > c) == up) {
            keys[lo] = c;
            keys[hi] = a;
          }
        }
        __syncthreads();
      }
    }
    // 5. next beam
    if (tid < beam) {
      const unsigned long long key = keys[tid];
      int2 bp = make_int2(-1, -1);
      if (key != ~0ull) {
        const int k = (int)(key & 0xffffull);
        Entry e;
        if (k < nbeam) {
          e = cur[k];
          e.b_prev = e.b_cur;
          e.nb_prev = e.nb_cur;
          bp = make_int2(k, -1);
        } else {
          const int ck = k - nbeam;
          const int i = ck / CW;
          const int c = cand_cls[ck];
          const float s = cand_logp[ck];
          e.b_prev = kNegInf;
          e.nb_prev = s;
          e.score = s;
          e.last = c;
          e.parent_id = cur[i].id;
          e.id = mix_id(cur[i].id, c);
          e.len = cur[i].len + 1;
          bp = make_int2(i, c);
        }
        e.b_cur = kNegInf;
        e.nb_cur = kNegInf;
        nxt[tid] = e;
      }
      backptr[row * beam + tid] = bp;
    }
    __syncthreads();
    if (tid == 0) {
      int n = 0;
      while (n < beam && keys[n] != ~0ull) ++n;
      s_nbeam = n;
    }
    if (tid < beam) cur[tid] = nxt[tid];
    __syncthreads();
  }

  // results: cur[] is already in prefix_compare order (score desc, last token asc) after the last step's sort;
  // for len == 0 it holds the root only.
  const int nbeam = s_nbeam;
  if (tid < beam) {
    int* o = ids + ((size_t)b * beam + tid) * T;
    if (tid < nbeam) {
      const Entry e = cur[tid];
      int n = e.len;
      for (int k = n; k < T; ++k) o[k] = -1;
      int slot = tid;
      for (int t = len - 1; t >= 0 && slot >= 0; --t) {
        const int2 bp = backptr[((size_t)b * T + t) * beam + slot];
        if (bp.y >= 0) o[--n] = bp.y;
        slot = bp.x;
      }
      out_len[b * beam + tid] = e.len;
      scores[b * beam + tid] = e.score;
    } else {
      for (int k = 0; k < T; ++k) o[k] = -1;
      out_len[b * beam + tid] = -1;
      scores[b * beam + tid] = -INFINITY;
    }
  }
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

// sorted classes kept per frame: a parent may have up to beam - 1 live children among its best classes, plus the blank, plus a repeat
// without blank mass, before its `beam` best NEW extensions are reached
static int beam_n_store(int V, int beam, int cutoff_top_n, float cutoff_prob) {
  int n = 2 * beam + 2;
  if (cutoff_prob < 1.0f) n = max(n, cutoff_top_n);
  return min(n, V);
}

size_t beam_workspace_bytes(int B, int T, int beam) {
  const size_t rows = (size_t)B * max(T, 1);
  return align256(rows * sizeof(FrameTop)) + 2 * align256(rows * kMaxTop * sizeof(float)) +
         align256(rows * beam * sizeof(int2));
}

int launch_ctc_beam(const BeamParams& p, cudaStream_t stream) {
  if (p.beam < 1 || p.beam > kMaxBeam) {
    snprintf(g_errbuf, sizeof(g_errbuf), "ctc_beam: beam size %d outside [1, %d]", p.beam, kMaxBeam);
    return 1;
  }
  if (p.cutoff_prob < 1.0f && (p.cutoff_top_n < 1 || p.cutoff_top_n > kMaxTop)) {
    snprintf(g_errbuf, sizeof(g_errbuf), "ctc_beam: cutoff_top_n %d outside [1, %d] (with cutoff_prob < 1)", p.cutoff_top_n, kMaxTop);
    return 1;
  }
  if (p.V < 2 || p.V > 65534 || p.blank < 0 || p.blank >= p.V) {
    snprintf(g_errbuf, sizeof(g_errbuf), "ctc_beam: bad vocabulary/blank (V=%d blank=%d)", p.V, p.blank);
    return 1;
  }
  if (p.B == 0) return 0;
  const int n_store = beam_n_store(p.V, p.beam, p.cutoff_top_n, p.cutoff_prob);
  const size_t rows = (size_t)p.B * max(p.T, 1);
  char* w = static_cast<char*>(p.workspace);
  FrameTop* meta = reinterpret_cast<FrameTop*>(w);
  w += align256(rows * sizeof(FrameTop));
  int* top_idx = reinterpret_cast<int*>(w);
  w += align256(rows * kMaxTop * sizeof(float));
  float* top_lp = reinterpret_cast<float*>(w);
  w += align256(rows * kMaxTop * sizeof(float));
  int2* backptr = reinterpret_cast<int2*>(w);
  if (p.T > 0) {
    const int nrows = p.B * p.T;
    beam_prep_kernel<<<ceil_div(nrows, 8), 256, 0, stream>>>(p.logits, nrows, p.V, n_store, p.cutoff_prob, p.cutoff_top_n, meta,
                                                             top_idx, top_lp, p.is_prob);
  }
  const int ncand = p.beam + p.beam * (p.beam + 1);
  if (ncand <= 64) {
    beam_search_kernel<64><<<p.B, 256, 0, stream>>>(p.logits, p.lengths, meta, top_idx, top_lp, p.T, p.V, p.blank, p.beam, n_store,
                                                    backptr, p.ids, p.out_len, p.scores, p.is_prob);
  } else if (ncand <= 512) {
    beam_search_kernel<512><<<p.B, 256, 0, stream>>>(p.logits, p.lengths, meta, top_idx, top_lp, p.T, p.V, p.blank, p.beam, n_store,
                                                     backptr, p.ids, p.out_len, p.scores, p.is_prob);
  } else {
    beam_search_kernel<2048><<<p.B, 256, 0, stream>>>(p.logits, p.lengths, meta, top_idx, top_lp, p.T, p.V, p.blank, p.beam,
                                                      n_store, backptr, p.ids, p.out_len, p.scores, p.is_prob);
  }
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200asr
