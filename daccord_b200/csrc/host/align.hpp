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
    if (m == 0) return n;
    if (m > 128) return align_dp(a, m, b, n, bmap);
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
        if (dv + dhup == cost) { bmap[i] = (uint32_t)j; --i; --j; }
        else if (dv == 1) { bmap[i] = (uint32_t)j; --i; }
        else --j;
      } else { bmap[i] = 0; --i; }
    }
    return score;
  }
  int align_dp(const uint8_t* a, int m, const uint8_t* b, int n, uint32_t* bmap) {
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
      if (j > 0 && cur == D[(size_t)(i - 1) * W + j - 1] + (a[i - 1] != b[j - 1])) { bmap[i] = (uint32_t)j; --i; --j; }
      else if (cur == D[(size_t)(i - 1) * W + j] + 1) { bmap[i] = (uint32_t)j; --i; }
      else --j;
    }
    return D[(size_t)m * W + n];
  }
};

}  // namespace dhost
