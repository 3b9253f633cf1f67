// TEST INFRASTRUCTURE ONLY (oracle).  Declaration-only stand-in for the reference's scorer.h (needs KenLM, which the
// reference downloads at build time).  Never instantiated: every call site is guarded by `ext_scorer != nullptr`.
#pragma once
#include <string>
#include <vector>
#include "path_trie.h"
class Scorer {
public:
  double alpha = 0, beta = 0;
  void* dictionary = nullptr;
  bool is_character_based() const { return true; }
  std::vector<std::string> make_ngram(PathTrie*) { return {}; }
  double get_log_cond_prob(const std::vector<std::string>&) { return 0; }
  double get_sent_log_prob(const std::vector<std::string>&) { return 0; }
  std::vector<std::string> split_labels(const std::vector<int>&) { return {}; }
};
