// TEST INFRASTRUCTURE ONLY (oracle): LOG(FATAL) as an aborting stream (decoder_utils.h:12-24 uses it for VALID_CHECK).
#pragma once
#include <cstdlib>
#include <iostream>
struct B200OracleFatal {
  ~B200OracleFatal() { std::cerr << std::endl; std::abort(); }
  template <class T> B200OracleFatal& operator<<(const T& v) { std::cerr << v; return *this; }
};
#define FATAL 0
#define LOG(x) B200OracleFatal()
