// simlas.hpp -- true overlaps between simulated reads, expressed as DALIGNER records with trace points
// (tspace-aligned A tiles, (diffs, blen) per tile) so that the real LAS -> trace -> window path can be exercised.
#pragma once
#include "synth.hpp"
#include "las.hpp"
#include <numeric>

namespace dhost {

// smallest genome offset o in [lo,hi] with g2fwd(o) >= fi
inline uint64_t fwd_to_goff(const SimRead& R, uint32_t fi, uint64_t lo, uint64_t hi) {
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (R.g2fwd(mid) >= fi) hi = mid; else lo = mid + 1; }
  return lo;
}

inline void make_overlaps(const std::vector<SimRead>& reads, int32_t tspace, uint64_t min_ovl, LasData& L) {
  L.tspace = tspace; L.ovl.clear(); L.trace.clear();
  const size_t n = reads.size();
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return reads[x].gstart != reads[y].gstart ? reads[x].gstart < reads[y].gstart : x < y; });
  std::vector<std::vector<uint32_t>> partners(n);
  for (size_t i = 0; i < n; ++i) {
    const SimRead& A = reads[order[i]];
    for (size_t j = i + 1; j < n; ++j) {
      const SimRead& B = reads[order[j]];
      if (B.gstart >= A.gstart + A.glen) break;
      uint64_t gs = std::max(A.gstart, B.gstart), ge = std::min(A.gstart + A.glen, B.gstart + B.glen);
      if (ge > gs && ge - gs >= min_ovl) { partners[order[i]].push_back(order[j]); partners[order[j]].push_back(order[i]); }
    }
  }
  for (size_t a = 0; a < n; ++a) {
    std::sort(partners[a].begin(), partners[a].end());
    const SimRead& A = reads[a]; const int64_t LA = (int64_t)A.seq.size();
    for (uint32_t b : partners[a]) {
      const SimRead& B = reads[b]; const int64_t LB = (int64_t)B.seq.size();
      uint64_t gs = std::max(A.gstart, B.gstart), ge = std::min(A.gstart + A.glen, B.gstart + B.glen);
      uint64_t alo = gs - A.gstart, ahi = ge - A.gstart;
      int64_t a0 = A.g2fwd(alo), a1 = A.g2fwd(ahi);
      if (a1 <= a0) continue;
      auto phi = [&](int64_t fi, uint64_t& goff_out) -> int64_t {       // A forward index -> B forward index
        uint64_t o = fwd_to_goff(A, (uint32_t)fi, alo, ahi);
        goff_out = o;
        return (int64_t)B.g2fwd(o + A.gstart - B.gstart);
      };
      Overlap O; O.aread = (int32_t)a; O.bread = (int32_t)b; O.flags = (A.rc != B.rc) ? 1u : 0u;
      O.abpos = (int32_t)(A.rc ? LA - a1 : a0); O.aepos = (int32_t)(A.rc ? LA - a0 : a1);
      O.trace_off = L.trace.size();
      int64_t x = O.abpos; int64_t totd = 0; int64_t bb = -1, be = -1;
      while (x < O.aepos) {
        int64_t y = std::min<int64_t>((x / tspace + 1) * tspace, O.aepos);
        uint64_t o1, o2; int64_t f1, f2;
        if (!A.rc) { f1 = phi(x, o1); f2 = phi(y, o2); } else { f2 = phi(LA - x, o2); f1 = phi(LA - y, o1); }   // f1 <= f2 in forward coordinates
        int64_t blen = f2 - f1;
        int64_t bs = A.rc ? LB - f2 : f1, bend = bs + blen;
        if (bb < 0) bb = bs;
        be = bend;
        int64_t d = (int64_t)(A.cerr[o2] - A.cerr[o1]) + (int64_t)(B.cerr[o2 + A.gstart - B.gstart] - B.cerr[o1 + A.gstart - B.gstart]);
        if (d > 255) d = 255;
        if (blen > 255 && tspace <= 125) blen = 255;    // cannot happen at sane error rates; keeps the byte encoding valid
        totd += d;
        L.trace.push_back((uint16_t)d); L.trace.push_back((uint16_t)blen);
        x = y;
      }
      O.tlen = (int32_t)(L.trace.size() - O.trace_off);
      O.diffs = (int32_t)totd; O.bbpos = (int32_t)bb; O.bepos = (int32_t)be;
      L.ovl.push_back(O);
    }
  }
  L.build_index(n);
}

}  // namespace dhost
