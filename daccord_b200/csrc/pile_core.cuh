// pile_core.cuh -- GPU side of the caller stage in front of the window kernel:
// trace reconstruction from DALIGNER trace points and window / slice extraction, i.e. daccord's
// OverlapDataInterface::computeTrace + advanceA / getStringLengthUsed bookkeeping and the active-set loop of
// HandleContext::operator() (reference src/HandleContext.hpp:1740-2049), producing the dcu_window / dcu_slice
// descriptors directly in HBM.  Three per-item routines, one GPU thread each:
//   pile_tile_starts : per overlap, B offset at which every trace tile starts           (prefix over the trace points)
//   pile_align_tile  : per tile, unit-cost global alignment of the A tile against its B block with 128-bit Myers
//                      vectors and a bit-vector traceback (rule: diagonal, DEL, INS); records the B offset reached
//                      after every A position that a window boundary can fall on
//   pile_read        : per A-read, the window loop: activation, expiry, order by (escore<<32)|z, slices
// Requirements of this path (checked by the host, which otherwise uses the host piler): w % a == 0, A tile <= 128.
// The same source is compiled for the host (tests/emu, -DDCU_EMU) to check it against the host piler without a GPU.
#pragma once
#include <stdint.h>

#ifdef DCU_EMU
#define PILE_FN static inline
#else
#define PILE_FN __device__ __forceinline__
#endif
#ifdef __CUDACC__
#define PILE_HD __host__ __device__ inline
#else
#define PILE_HD static inline
#endif

namespace dpile {

struct Ovl {                     // one selected overlap (host order: by A-read, then abpos, ties in file order)
  int32_t abpos, aepos, bbpos, bread;
  uint32_t flags; int32_t aread; int32_t diffs; int32_t ntiles;
  uint64_t trace_off;            // index of the first (diffs, blen) pair value in the trace array
  uint64_t tile_off;             // first tile index of this overlap
  uint64_t bm_off;               // first entry of this overlap in the boundary offset array
  uint32_t ridx, pad;            // index of the A-read in the batch
};
struct ReadInfo { uint64_t ovl_begin, ovl_end; uint64_t win_off, sl_off; uint32_t maxaepos; uint32_t nwin, nsl; };
struct Params { int32_t tspace; uint32_t w, a; uint64_t maxalign; };

PILE_FN uint8_t base_at(const uint8_t* packed, uint64_t boff, uint32_t len, uint32_t pos, bool comp) {
  uint32_t g = comp ? len - 1 - pos : pos;
  uint8_t b = (packed[boff + (g >> 2)] >> (6 - 2 * (g & 3))) & 3;
  return comp ? (uint8_t)(3 - b) : b;
}

// number of boundary entries of an overlap: A positions p in [abpos, aepos] with p % a == 0, plus two specials
PILE_HD uint32_t bm_entries(int32_t abpos, int32_t aepos, uint32_t a) {
  int64_t first = ((int64_t)abpos + a - 1) / a, last = (int64_t)aepos / a;
  return (uint32_t)(last >= first ? last - first + 1 : 0) + 2;
}
PILE_FN uint32_t bm_index(int32_t abpos, uint32_t a, uint32_t p) { return (uint32_t)(p / a - ((uint32_t)abpos + a - 1) / a); }

PILE_FN void pile_tile_starts(const Ovl& o, const uint16_t* trace, uint32_t* tile_b) {
  uint32_t b = 0;
  for (int32_t t = 0; t < o.ntiles; ++t) { tile_b[o.tile_off + t] = b; b += trace[o.trace_off + 2 * t + 1]; }
}

struct U128 { unsigned long long lo, hi; };
PILE_FN U128 u_and(U128 a, U128 b) { return {a.lo & b.lo, a.hi & b.hi}; }
PILE_FN U128 u_or(U128 a, U128 b) { return {a.lo | b.lo, a.hi | b.hi}; }
PILE_FN U128 u_xor(U128 a, U128 b) { return {a.lo ^ b.lo, a.hi ^ b.hi}; }
PILE_FN U128 u_not(U128 a) { return {~a.lo, ~a.hi}; }
PILE_FN U128 u_add(U128 a, U128 b) { U128 r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull); return r; }
PILE_FN U128 u_shl1(U128 a) { return {a.lo << 1, (a.hi << 1) | (a.lo >> 63)}; }
PILE_FN int u_bit(U128 a, int i) { return (int)(((i < 64 ? a.lo >> i : a.hi >> (i - 64))) & 1ull); }

enum { PILE_MAXB = 256 };

// one tile of one overlap.  special0 / special1: absolute A positions (l - w and l of the read) whose offsets are also recorded
PILE_FN void pile_align_tile(const Ovl& o, int tile, const Params& P, const uint16_t* trace, const uint32_t* tile_b, const uint8_t* packed,
                             const uint64_t* read_boff, const uint32_t* read_len, uint32_t special0, uint32_t special1, uint32_t* bm,
                             U128* PV, U128* MV, U128* PH, U128* MH /* PILE_MAXB+1 entries each, thread private */) {
  const int64_t x0 = o.abpos;
  const int64_t x = tile == 0 ? x0 : ((x0 / P.tspace) + tile) * (int64_t)P.tspace;
  int64_t y = ((x / P.tspace) + 1) * (int64_t)P.tspace; if (y > o.aepos) y = o.aepos;
  const int m = (int)(y - x);
  const int n = (int)trace[o.trace_off + 2 * tile + 1];
  const uint32_t bstart = (uint32_t)o.bbpos + tile_b[o.tile_off + tile];
  const bool comp = (o.flags & 1u) != 0;
  const uint64_t aoff = read_boff[o.aread], boff = read_boff[o.bread];
  const uint32_t alen = read_len[o.aread], blen = read_len[o.bread];
  const uint32_t base = tile_b[o.tile_off + tile];
  const uint32_t nmul = bm_entries(o.abpos, o.aepos, P.a) - 2;
  uint32_t* out = bm + o.bm_off;
  if (tile == 0) {
    if ((uint32_t)x0 % P.a == 0) out[bm_index(o.abpos, P.a, (uint32_t)x0)] = 0;
    if (special0 == (uint32_t)x0) out[nmul] = 0;
    if (special1 == (uint32_t)x0) out[nmul + 1] = 0;
  }
  if (m <= 0) return;
  U128 peq[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  for (int i = 0; i < m; ++i) { int ch = base_at(packed, aoff, alen, (uint32_t)(x + i), false); if (i < 64) peq[ch].lo |= 1ull << i; else peq[ch].hi |= 1ull << (i - 64); }
  U128 pv = {~0ull, ~0ull}, mv = {0, 0};
  for (int j = 1; j <= n; ++j) {
    U128 eq = peq[base_at(packed, boff, blen, bstart + (uint32_t)(j - 1), comp)];
    U128 xv = u_or(eq, mv);
    U128 xh = u_or(u_xor(u_add(u_and(eq, pv), pv), pv), eq);
    U128 ph = u_or(mv, u_not(u_or(xh, pv)));
    U128 mh = u_and(pv, xh);
    PH[j] = ph; MH[j] = mh;
    ph = u_shl1(ph); ph.lo |= 1ull; mh = u_shl1(mh);
    pv = u_or(mh, u_not(u_or(xv, ph))); mv = u_and(ph, xv);
    PV[j] = pv; MV[j] = mv;
  }
  int i = m, j = n;
  while (i > 0) {
    int step;     // 0 diagonal, 1 DEL (consume A), 2 INS (consume B)
    if (j > 0) {
      int dv = u_bit(PV[j], i - 1) ? 1 : (u_bit(MV[j], i - 1) ? -1 : 0);
      int dhup = (i == 1) ? 1 : (u_bit(PH[j], i - 2) ? 1 : (u_bit(MH[j], i - 2) ? -1 : 0));
      int cost = base_at(packed, aoff, alen, (uint32_t)(x + i - 1), false) != base_at(packed, boff, blen, bstart + (uint32_t)(j - 1), comp);
      step = (dv + dhup == cost) ? 0 : (dv == 1 ? 1 : 2);
    } else step = 1;
    if (step == 2) { --j; continue; }
    // the i-th A symbol of the tile is consumed by this step, which ends in column j
    const uint32_t p = (uint32_t)(x + i), v = base + (uint32_t)j;
    if (p % P.a == 0) out[bm_index(o.abpos, P.a, p)] = v;
    if (p == special0) out[nmul] = v;
    if (p == special1) out[nmul + 1] = v;
    --i; if (step == 0) --j;
  }
}

// B offset (relative to bbpos) once exactly p - abpos A symbols are consumed; p must be a recorded boundary
PILE_FN uint32_t bm_lookup(const Ovl& o, const Params& P, const uint32_t* bm, uint32_t p, uint32_t special0, uint32_t special1) {
  const uint32_t nmul = bm_entries(o.abpos, o.aepos, P.a) - 2;
  if (p % P.a == 0) return bm[o.bm_off + bm_index(o.abpos, P.a, p)];
  return bm[o.bm_off + nmul + (p == special0 ? 0 : 1)];
  (void)special1;
}

struct Win { uint32_t slice_begin; uint16_t slice_cnt; uint16_t reserved; uint32_t aread; uint32_t astart; };
struct Sl { uint32_t gpos; uint16_t len; uint16_t flags; };

// windows of HandleContext::Windows (reference src/HandleContext.hpp:382-447)
PILE_FN uint64_t win_count(uint64_t l, uint64_t a, uint64_t w) {
  uint64_t npre = (l + a >= w) ? ((l + a - w) / a) : 0;
  if (npre) return ((npre - 1) * a + w == l) ? npre : npre + 1;
  return l >= w ? 1 : 0;
}
PILE_FN uint64_t win_start(uint64_t i, uint64_t l, uint64_t a, uint64_t w) { return (i * a + w <= l) ? i * a : l - w; }

// The window loop of one A-read.  fill == false: count windows and slices only.  `act` is this read's scratch list of
// active overlaps (key = (escore << 32) | z, kept sorted), at most cap entries.  Returns 1 on a capacity problem.
PILE_FN int pile_read(const ReadInfo& R, const Ovl* ovl, const Params& P, const uint32_t* bm, const uint64_t* read_boff, const uint32_t* read_len,
                      double minerate, double ediv, bool fill, Win* win, Sl* sl, uint32_t* nwin_out, uint32_t* nsl_out,
                      unsigned long long* act, int cap, uint32_t aread) {
  const uint64_t nintv = R.ovl_end - R.ovl_begin;
  uint32_t nw = 0, ns = 0;
  if (!nintv) { *nwin_out = 0; *nsl_out = 0; return 0; }
  const uint64_t l = R.maxaepos;
  const uint64_t W = win_count(l, P.a, P.w);
  const uint32_t special0 = l >= P.w ? (uint32_t)(l - P.w) : 0, special1 = (uint32_t)l;
  int nact = 0; uint64_t z = 0;
  for (uint64_t y = 0; y < W; ++y) {
    const uint64_t astart = win_start(y, l, P.a, P.w), aend = astart + P.w;
    while (z < nintv && (int64_t)astart >= ovl[R.ovl_begin + z].abpos) {             // activation (:1904-1967)
      const Ovl& o = ovl[R.ovl_begin + z];
      if (o.aepos >= (int64_t)aend) {
        double er = (double)o.diffs / (double)(o.aepos - o.abpos);
        unsigned long long escore = (unsigned long long)(((er - minerate) / ediv) * 4294967295.0);
        unsigned long long key = (escore << 32) | z;
        if (nact >= cap) return 1;
        int q = nact++;
        while (q > 0 && act[q - 1] > key) { act[q] = act[q - 1]; --q; }
        act[q] = key;
      }
      ++z;
    }
    { int q = 0;                                                                     // expiry (:1969-1977)
      for (int t = 0; t < nact; ++t) { const Ovl& o = ovl[R.ovl_begin + (act[t] & 0xFFFFFFFFull)]; if (!((uint64_t)o.aepos < aend)) act[q++] = act[t]; }
      nact = q; }
    uint64_t MAo = 0; const uint32_t sbegin = ns;
    for (int t = 0; t < nact; ++t) {                                                 // slices (:1984-2049)
      const Ovl& o = ovl[R.ovl_begin + (act[t] & 0xFFFFFFFFull)];
      if (!MAo) { if (fill) { sl[R.sl_off + ns].gpos = (uint32_t)(read_boff[aread] * 4 + astart); sl[R.sl_off + ns].len = (uint16_t)P.w; sl[R.sl_off + ns].flags = 0; } ++ns; ++MAo; }
      if (MAo < P.maxalign) {
        if (fill) {
          uint32_t b0 = bm_lookup(o, P, bm, (uint32_t)astart, special0, special1), b1 = bm_lookup(o, P, bm, (uint32_t)aend, special0, special1);
          uint32_t s = (uint32_t)o.bbpos + b0, len = b1 - b0, LB = read_len[o.bread];
          bool comp = (o.flags & 1u) != 0;
          uint64_t g = read_boff[o.bread] * 4 + (comp ? (uint64_t)(LB - s - len) : (uint64_t)s);
          if (len > 255) return 1;
          sl[R.sl_off + ns].gpos = (uint32_t)g; sl[R.sl_off + ns].len = (uint16_t)len; sl[R.sl_off + ns].flags = (uint16_t)(comp ? 1 : 0);
        }
        ++ns; ++MAo;
      }
    }
    if (MAo) {
      if (fill) { Win& wv = win[R.win_off + nw]; wv.slice_begin = (uint32_t)(R.sl_off + sbegin); wv.slice_cnt = (uint16_t)(MAo > 65535 ? 65535 : MAo); wv.reserved = 0; wv.aread = aread; wv.astart = (uint32_t)astart; }
      ++nw;
    }
  }
  *nwin_out = nw; *nsl_out = ns;
  return 0;
}

}  // namespace dpile
