// align.hpp -- host-side unit-cost global alignment of one trace tile (A tile vs B block) with the traceback
// rule this repository fixes for libmaus2's aligner (diagonal, then DEL = consume A, then INS = consume B;
// reference call site OverlapDataInterface::computeTrace, src/HandleContext.hpp:1914).  Bit-vector (Myers)
// over 128-bit words, column vectors kept for the traceback; plain DP for tiles longer than 128.
// Instead of materialising the step string it returns what the window extractor needs from it
// (AlignmentTraceContainer::advanceA / getStringLengthUsed, src/HandleContext.hpp:1936-1949, :2005-2029):
// bmap[i] = number of B symbols consumed when exactly i A symbols have been consumed.
#pragma once
#include <cstdint>
#include <vector>
#include <algorithm>

namespace dhost {

typedef unsigned __int128 u128;

struct TileAligner {
  std::vector<u128> PV, MV, PH, MH;
  std::vector<int32_t> D;
  // a: m symbols (codes 0..3), b: n symbols; fills bmap[0..m]; returns edit distance
  int align(const uint8_t* a, int m, const uint8_t* b, int n, uint32_t* bmap) {
    bmap[0] = 0;
    return walk(a, m, b, n, [&](int kind, int i, int j) { if (kind != 2) bmap[i] = (uint32_t)j; });
  }
  // same alignment, counted: cnt[0] matches, [1] mismatches, [2] insertions (B symbol only), [3] deletions (A symbol only);
  // what libmaus2's AlignmentStatistics holds after Aligner::align(a, b) (reference call site src/daccord.cpp:588-597)
  int align_count(const uint8_t* a, int m, const uint8_t* b, int n, uint64_t cnt[4]) {
    return walk(a, m, b, n, [&](int kind, int i, int j) { if (kind == 0) cnt[a[i - 1] != b[j - 1]]++; else cnt[kind == 1 ? 3 : 2]++; });
  }

 private:
  // forward pass + traceback; emit(kind, i, j) is called back to front, kind 0 = diagonal consuming a[i-1], b[j-1],
  // 1 = DEL consuming a[i-1] with j B symbols before it, 2 = INS consuming b[j-1]
  template <class F> int walk(const uint8_t* a, int m, const uint8_t* b, int n, F emit) {
    if (m == 0) { for (int j = n; j > 0; --j) emit(2, 0, j); return n; }
    if (m > 128) return walk_dp(a, m, b, n, emit);
    u128 peq[4] = {0, 0, 0, 0};
    for (int i = 0; i < m; ++i) peq[a[i] & 3] |= (u128)1 << i;
    PV.resize(n + 1); MV.resize(n + 1); PH.resize(n + 1); MH.resize(n + 1);
    u128 pv = ~(u128)0, mv = 0; const u128 top = (u128)1 << (m - 1);
    int score = m;
    for (int j = 1; j <= n; ++j) {
      u128 eq = peq[b[j - 1] & 3];
      u128 xv = eq | mv;
      u128 xh = (((eq & pv) + pv) ^ pv) | eq;
      u128 ph = mv | ~(xh | pv);
      u128 mh = pv & xh;
      if (ph & top) ++score; else if (mh & top) --score;
      PH[j] = ph; MH[j] = mh;
      ph = (ph << 1) | (u128)1; mh <<= 1;
      pv = mh | ~(xv | ph); mv = ph & xv;
      PV[j] = pv; MV[j] = mv;
    }
    int i = m, j = n;
    while (i > 0) {
      if (j > 0) {
        int dv = ((PV[j] >> (i - 1)) & 1) ? 1 : (((MV[j] >> (i - 1)) & 1) ? -1 : 0);
        int dhup = (i == 1) ? 1 : (((PH[j] >> (i - 2)) & 1) ? 1 : (((MH[j] >> (i - 2)) & 1) ? -1 : 0));
        int cost = a[i - 1] != b[j - 1];
        if (dv + dhup == cost) { emit(0, i, j); --i; --j; }
        else if (dv == 1) { emit(1, i, j); --i; }
        else { emit(2, i, j); --j; }
      } else { emit(1, i, 0); --i; }
    }
    for (; j > 0; --j) emit(2, 0, j);
    return score;
  }
  template <class F> int walk_dp(const uint8_t* a, int m, const uint8_t* b, int n, F emit) {
    const int W = n + 1;
    D.assign((size_t)(m + 1) * W, 0);
    for (int j = 0; j <= n; ++j) D[j] = j;
    for (int i = 1; i <= m; ++i) {
      D[(size_t)i * W] = i;
      for (int j = 1; j <= n; ++j) D[(size_t)i * W + j] = std::min(D[(size_t)(i - 1) * W + j - 1] + (a[i - 1] != b[j - 1]), std::min(D[(size_t)(i - 1) * W + j] + 1, D[(size_t)i * W + j - 1] + 1));
    }
    int i = m, j = n;
    while (i > 0) {
      int cur = D[(size_t)i * W + j];
      if (j > 0 && cur == D[(size_t)(i - 1) * W + j - 1] + (a[i - 1] != b[j - 1])) { emit(0, i, j); --i; --j; }
      else if (cur == D[(size_t)(i - 1) * W + j] + 1) { emit(1, i, j); --i; }
      else { emit(2, i, j); --j; }
    }
    for (; j > 0; --j) emit(2, 0, j);
    return D[(size_t)m * W + n];
  }
};

}  // namespace dhost
