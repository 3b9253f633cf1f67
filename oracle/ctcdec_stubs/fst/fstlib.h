// TEST INFRASTRUCTURE ONLY (oracle).  Minimal stand-in for <fst/fstlib.h> (openfst-1.6.3 is fetched by the reference's
// externals/ctc_decoders/setup.sh, not vendored).  Only the declarations path_trie.{h,cpp} / decoder_utils.{h,cpp} /
// ctc_beam_search_decoder.cpp need to *compile*; with ext_scorer == nullptr none of this is ever executed.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace fst {
struct TropicalWeight {
  float v;
  TropicalWeight(float x = 0.f) : v(x) {}
  static TropicalWeight Zero() { return TropicalWeight(std::numeric_limits<float>::infinity()); }
  static TropicalWeight One() { return TropicalWeight(0.f); }
  bool operator!=(const TropicalWeight& o) const { return v != o.v; }
};
struct StdArc {
  typedef TropicalWeight Weight;
  typedef int StateId;
  int ilabel, olabel;
  Weight weight;
  StateId nextstate;
  StdArc(int i = 0, int o = 0, Weight w = Weight(), StateId n = 0) : ilabel(i), olabel(o), weight(w), nextstate(n) {}
};
struct StdVectorFst {
  typedef int StateId;
  std::vector<std::vector<StdArc>> states;
  std::vector<TropicalWeight> finals;
  StateId start = -1;
  StateId NumStates() const { return (StateId)states.size(); }
  StateId AddState() { states.emplace_back(); finals.push_back(TropicalWeight::Zero()); return (StateId)states.size() - 1; }
  void SetStart(StateId s) { start = s; }
  StateId Start() const { return start; }
  void AddArc(StateId s, const StdArc& a) { states[s].push_back(a); }
  void SetFinal(StateId s, TropicalWeight w) { finals[s] = w; }
  TropicalWeight Final(StateId s) const { return finals[s]; }
  StdVectorFst* Copy(bool = false) const { return new StdVectorFst(*this); }
};
enum MatchType { MATCH_INPUT = 1 };
template <class F>
struct SortedMatcher {
  const F& f;
  int state = 0;
  StdArc cur;
  SortedMatcher(const F& fst_, MatchType) : f(fst_) {}
  void SetState(int s) { state = s; }
  bool Find(int label) {
    for (const auto& a : f.states[state])
      if (a.ilabel == label) { cur = a; return true; }
    return false;
  }
  const StdArc& Value() const { return cur; }
};
}  // namespace fst
