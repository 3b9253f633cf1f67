// TEST INFRASTRUCTURE ONLY (oracle).  C entry points over the reference's own externals/ctc_decoders C++ (compiled from
// the zip by oracle/build_ref.py) so tests can call ctc_beam_search_decoder / ctc_greedy_decoder with integer ids.
// The vocabulary handed to the reference is the decimal id of each class followed by ',' (no spaces -> the word-
// timestamp post-processing in decoder_utils.cpp:71-85 stays inert), so output strings parse back into ids.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ctc_beam_search_decoder.h"
#include "ctc_greedy_decoder.h"

static std::vector<std::string> make_vocab(int n) {
  std::vector<std::string> v;
  for (int i = 0; i < n; ++i) v.push_back(std::to_string(i) + ",");
  return v;
}

static int parse_ids(const std::string& s, int* out, int cap) {
  int n = 0;
  size_t pos = 0;
  while (pos < s.size()) {
    size_t e = s.find(',', pos);
    if (e == std::string::npos) break;
    if (n < cap) out[n] = atoi(s.substr(pos, e - pos).c_str());
    ++n;
    pos = e + 1;
  }
  return n;
}

extern "C" {

// probs [T, V] row-major doubles (blank = V-1).  Outputs up to `beam` hypotheses: ids [beam, T], lens [beam], scores [beam].
// Returns the number of hypotheses.
int ctcref_beam(const double* probs, int T, int V, int beam, double cutoff_prob, int cutoff_top_n, int* ids, int* lens,
                double* scores) {
  std::vector<std::vector<double>> seq(T, std::vector<double>(V));
  for (int t = 0; t < T; ++t) memcpy(seq[t].data(), probs + (size_t)t * V, sizeof(double) * V);
  auto vocab = make_vocab(V - 1);
  auto res = ctc_beam_search_decoder(seq, vocab, (size_t)beam, cutoff_prob, (size_t)cutoff_top_n, nullptr);
  int n = 0;
  for (auto& r : res) {
    if (n >= beam) break;
    scores[n] = r.first;
    lens[n] = parse_ids(r.second, ids + (size_t)n * T, T);
    ++n;
  }
  return n;
}

int ctcref_greedy(const double* probs, int T, int V, int* ids) {
  std::vector<std::vector<double>> seq(T, std::vector<double>(V));
  for (int t = 0; t < T; ++t) memcpy(seq[t].data(), probs + (size_t)t * V, sizeof(double) * V);
  auto vocab = make_vocab(V - 1);
  std::string s = ctc_greedy_decoder(seq, vocab);
  return parse_ids(s, ids, T);
}

}  // extern "C"
