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

// edit distance only (two-row DP)
inline uint64_t editDistance(const uint8_t* a, uint64_t la, const uint8_t* b, uint64_t lb) {
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
