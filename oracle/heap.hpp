// TEST INFRASTRUCTURE (oracle) -- parity unpinned, see oracle/README.md.
// Bounded binary heap standing in for libmaus2::util::FiniteSizeHeap (not in
// /root/reference; call sites e.g. src/DebruijnGraph.hpp:3602-3665, :4851-4881, :5063-5106,
// src/HandleContext.hpp:1962-1977).  Convention C2: top() is the least element under Cmp.
#pragma once
#include <vector>
#include <cstdint>
#include <utility>
#include <functional>

namespace oracle {

template <typename T, typename Cmp = std::less<T>>
struct FiniteHeap {
  std::vector<T> H;
  uint64_t f = 0;
  uint64_t cap;
  Cmp cmp;
  explicit FiniteHeap(uint64_t c = 0, Cmp rc = Cmp()) : H(c), cap(c), cmp(rc) {}
  bool empty() const { return f == 0; }
  bool full() const { return f == cap; }
  void clear() { f = 0; }
  const T& top() const { return H[0]; }
  void push(const T& x) {                       // requires !full()
    uint64_t i = f++;
    H[i] = x;
    while (i > 0) {
      uint64_t p = (i - 1) >> 1;
      if (cmp(H[i], H[p])) { std::swap(H[i], H[p]); i = p; } else break;
    }
  }
  void pushBump(const T& x) {                   // push that grows the capacity
    if (f == cap) { cap = cap ? 2 * cap : 1; H.resize(cap); }
    push(x);
  }
  void popvoid() {
    H[0] = H[--f];
    uint64_t p = 0;
    for (;;) {
      uint64_t l = 2 * p + 1, r = l + 1;
      if (l >= f) break;
      uint64_t m = (r < f && cmp(H[r], H[l])) ? r : l;
      if (cmp(H[m], H[p])) { std::swap(H[m], H[p]); p = m; } else break;
    }
  }
  T pop() { T t = H[0]; popvoid(); return t; }
};

}  // namespace oracle
