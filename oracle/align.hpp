// TEST INFRASTRUCTURE (oracle) -- parity unpinned, see oracle/README.md.
// Unit-cost global alignment standing in for libmaus2::lcs::Aligner (NP / SIMD banded) and
// libmaus2::lcs::AlignmentOneAgainstMany; call sites: src/HandleContext.hpp:1914, :2434,
// src/DebruijnGraph.hpp:5361, :5430.  Convention C1 (traceback: diagonal, then DEL, then INS).
#pragma once
#include <cstdint>
#include <vector>
#include <algorithm>

namespace oracle {

enum Step : uint8_t { STEP_MATCH = 0, STEP_MISMATCH = 1, STEP_INS = 2, STEP_DEL = 3 };
// STEP_INS consumes a b symbol only, STEP_DEL an a symbol only (HandleContext.hpp:2451-2491)

struct Aligner {
  std::vector<int32_t> D;
  std::vector<uint8_t> trace;   // forward order
  // global alignment of a[0,la) vs b[0,lb); returns edit distance, fills trace
  uint64_t align(const uint8_t* a, uint64_t la, const uint8_t* b, uint64_t lb) {
    uint64_t W = lb + 1;
    D.assign((la + 1) * W, 0);
    for (uint64_t j = 0; j <= lb; ++j) D[j] = (int32_t)j;
    for (uint64_t i = 1; i <= la; ++i) {
      D[i * W] = (int32_t)i;
      for (uint64_t j = 1; j <= lb; ++j) {
        int32_t d = D[(i - 1) * W + (j - 1)] + (a[i - 1] != b[j - 1]);
        int32_t u = D[(i - 1) * W + j] + 1;
        int32_t l = D[i * W + (j - 1)] + 1;
        D[i * W + j] = std::min(d, std::min(u, l));
      }
    }
    trace.clear();
    uint64_t i = la, j = lb;
    while (i || j) {
      int32_t c = D[i * W + j];
      if (i && j && c == D[(i - 1) * W + (j - 1)] + (a[i - 1] != b[j - 1])) {
        trace.push_back(a[i - 1] != b[j - 1] ? STEP_MISMATCH : STEP_MATCH); --i; --j;
      } else if (i && c == D[(i - 1) * W + j] + 1) {
        trace.push_back(STEP_DEL); --i;
      } else {
        trace.push_back(STEP_INS); --j;
      }
    }
    std::reverse(trace.begin(), trace.end());
    return (uint64_t)D[la * W + lb];
  }
};

// edit distance only (two-row DP): the definition; kept as the check of the bit-parallel version below (tests/test_cpu_parity.py)
inline uint64_t editDistanceDP(const uint8_t* a, uint64_t la, const uint8_t* b, uint64_t lb) {
  std::vector<int32_t> prev(lb + 1), cur(lb + 1);
  for (uint64_t j = 0; j <= lb; ++j) prev[j] = (int32_t)j;
  for (uint64_t i = 1; i <= la; ++i) {
    cur[0] = (int32_t)i;
    for (uint64_t j = 1; j <= lb; ++j)
      cur[j] = std::min(prev[j - 1] + (a[i - 1] != b[j - 1]), std::min(prev[j] + 1, cur[j - 1] + 1));
    std::swap(prev, cur);
  }
  return (uint64_t)prev[lb];
}
// Edit distance as the reference computes it for candidate scoring: libmaus2 runs a bit-parallel / SIMD aligner there
// (src/DebruijnGraphBase.hpp:26-47), so a full int32 DP would make the CPU arm of the bench slower than real daccord.
// Myers' bit-vector algorithm (one 64-bit word: the shorter string is the pattern; the distance is symmetric), value
// identical to the DP -- no convention is involved, only the number.  Strings longer than 64 on both sides take the DP.
inline uint64_t editDistance(const uint8_t* a, uint64_t la, const uint8_t* b, uint64_t lb) {
  if (la > lb) { std::swap(a, b); std::swap(la, lb); }
  if (la == 0) return lb;
  if (la > 64) return editDistanceDP(a, la, b, lb);
  uint64_t peq[256] = {0};            // only the symbols of a are ever set; cleared again below
  for (uint64_t i = 0; i < la; ++i) peq[a[i]] |= 1ull << i;
  uint64_t pv = ~0ull, mv = 0; const uint64_t top = 1ull << (la - 1);
  int64_t score = (int64_t)la;
  for (uint64_t j = 0; j < lb; ++j) {
    const uint64_t eq = peq[b[j]];
    const uint64_t xv = eq | mv;
    const uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
    uint64_t ph = mv | ~(xh | pv);
    uint64_t mh = pv & xh;
    if (ph & top) ++score; else if (mh & top) --score;
    ph = (ph << 1) | 1ull; mh <<= 1;
    pv = mh | ~(xv | ph); mv = ph & xv;
  }
  return (uint64_t)score;
}

// AlignmentTraceContainer::advanceA (call sites src/HandleContext.hpp:1936, :2005, :2023):
// consume steps until n a-symbols are covered; returns (a covered, steps consumed)
inline std::pair<uint64_t, uint64_t> advanceA(const uint8_t* ta, const uint8_t* te, uint64_t n) {
  const uint8_t* tc = ta; uint64_t a = 0;
  while (tc != te && a < n) {
    switch (*tc++) { case STEP_MATCH: case STEP_MISMATCH: case STEP_DEL: ++a; break; default: break; }
  }
  return {a, (uint64_t)(tc - ta)};
}
// AlignmentTraceContainer::getStringLengthUsed (src/HandleContext.hpp:1946, :2018)
inline std::pair<uint64_t, uint64_t> stringLengthUsed(const uint8_t* ta, const uint8_t* te) {
  uint64_t a = 0, b = 0;
  for (; ta != te; ++ta) switch (*ta) {
    case STEP_MATCH: case STEP_MISMATCH: ++a; ++b; break;
    case STEP_INS: ++b; break;
    case STEP_DEL: ++a; break;
  }
  return {a, b};
}

}  // namespace oracle
