// vote_core.cuh -- GPU side of the caller stage behind the window kernel: the pile vote of
// HandleContext::operator() (reference src/HandleContext.hpp:2446-2493 placement walk, :2541-2706 sort + column vote),
// computed straight from the resident results so that corrected bases instead of per-window traces leave the GPU.
//
// The reference pushes one PileElement (apos, apre, sym) per consensus symbol of every window, sorts them and votes
// per (apos, apre) column from right to left.  Here no element is materialised and nothing is sorted:
//   vote_window_table : per window, one 16-bit entry per A offset 0..w: where in the consensus the symbol placed on that
//                       A base sits (or that the base is deleted) and how many consensus symbols are inserted before it
//   vote_position     : per A position, gathers the <= w/a + 2 windows covering it and votes its columns -- the
//                       insertion columns (apre = -q) from the windows that insert at least q symbols there, the base
//                       column (apre = 0) from all of them -- with the reference's rules: missing members of an
//                       insertion column count as 'D' up to the depth of the nearest base column to the right,
//                       ties go to the larger (count, character) pair, 'D' winners emit nothing
// Both passes of vote_position (count, then write at the scanned offsets) run the same code.  Runs of consecutive
// positions (:2590-2612) are reported as boundary records; pairing and the length filters are host work (vote_host.hpp).
// The same source is compiled for the host (tests/emu, -DDCU_EMU) and checked against host/vote.hpp without a GPU.
#pragma once
#include <stdint.h>

#ifdef DCU_EMU
#define VOTE_FN static inline
#else
#define VOTE_FN __device__ __forceinline__
#endif

namespace dvote {

struct Win { uint32_t slice_begin; uint16_t slice_cnt; uint16_t reserved; uint32_t aread; uint32_t astart; };            // = dcu_window
struct Res { uint8_t status, k; int8_t ff; uint8_t clen; uint32_t err; uint16_t nops, ncand; int32_t elength; };         // = dcu_result
struct Read {                    // one A-read of the batch
  uint64_t pos_off;              // index of its position 0 in the per-position arrays
  uint64_t boff;                 // byte offset in the packed database (-f only)
  uint32_t aread, wb, we;        // its windows [wb, we), ascending astart
  uint32_t span;                 // positions 0 .. span-1 are examined
  uint32_t rlen, pad;            // read length (-f only, else 0)
};
struct Bound { uint64_t off; uint32_t ridx, pos; uint32_t kind, pad; };   // kind 0: run starts at pos (off = first char), 1: run ends at pos (off = one past its last char)
struct Params { uint32_t w; uint32_t cons_stride, ops_stride; int32_t producefull; };

enum { ST_OK = 1, OP_MATCH = 0, OP_MISMATCH = 1, OP_INS = 2, OP_DEL = 3 };
enum { E_DEL = 0x8000 };         // entry = DEL flag | (consensus index after the insertions) << 8 | insertions before this base

// placement walk of one window (:2446-2493) folded into w + 1 entries; false if the trace does not cover exactly w A bases
VOTE_FN bool vote_window_table(const Res& R, const uint8_t* ops, uint32_t w, uint16_t* ent) {
  uint32_t c = 0, t = 0, j = 0;
  const uint32_t nops = R.nops;
  uint32_t trailing = 0;
  while (t < nops) {
    uint32_t nins = 0;
    while (t < nops && ops[t] == OP_INS) { ++nins; ++t; }
    c += nins;
    if (t < nops) {
      const uint8_t op = ops[t++];
      if (j >= w) return false;
      if (op == OP_DEL) ent[j] = (uint16_t)(E_DEL | (c << 8) | nins);
      else { ent[j] = (uint16_t)((c << 8) | nins); ++c; }
      ++j;
    } else trailing = nins;
  }
  if (j != w || c != R.clen) return false;
  ent[w] = (uint16_t)((c << 8) | trailing);
  return true;
}

struct Ctx {
  const Win* win; const Res* res; const uint8_t* cons; const uint16_t* ent; const uint8_t* packed;
  Params P;
};

// first window of the read whose range [astart, astart + w] can contain p
VOTE_FN uint32_t first_cover(const Ctx& c, const Read& R, uint32_t p) {
  if (p <= c.P.w) return R.wb;
  const uint32_t lim = p - c.P.w;
  uint32_t a = R.wb, b = R.we;
  while (a < b) { uint32_t mid = (a + b) >> 1; if (c.win[mid].astart < lim) a = mid + 1; else b = mid; }
  return a;
}
// what the pile holds at position p: number of base-column members, longest insertion; returns whether anything is there
VOTE_FN bool scan_pos(const Ctx& c, const Read& R, uint32_t p, int* ld0, int* maxq, bool* filled) {
  int l0 = 0, mq = 0; bool present = false;
  const uint32_t w = c.P.w;
  for (uint32_t i = first_cover(c, R, p); i < R.we && c.win[i].astart <= p; ++i) {
    if (c.res[i].status != ST_OK) continue;
    const uint32_t j = p - c.win[i].astart;
    const uint32_t e = c.ent[(uint64_t)i * (w + 1) + j];
    const int ni = (int)(e & 0xFF);
    if (j < w) { ++l0; present = true; if (ni > mq) mq = ni; }
    else if (ni) { present = true; if (ni > mq) mq = ni; }
  }
  *filled = false;
  if (!present && c.P.producefull && p < R.rlen) { present = true; l0 = 1; *filled = true; }     // :2543-2580
  *ld0 = l0; *maxq = mq;
  return present;
}
// larger (count, character) pair wins (:2690 std::sort with std::greater); cnt order A C G T D
VOTE_FN int pick(const int cnt[5]) {
  const char sym[5] = {'A', 'C', 'G', 'T', 'D'};
  int best = -1;
  for (int s = 0; s < 5; ++s)
    if (best < 0 || cnt[s] > cnt[best] || (cnt[s] == cnt[best] && sym[s] > sym[best])) best = s;
  return (cnt[best] > 0 && best != 4) ? sym[best] : 0;
}
VOTE_FN int sym_index(uint8_t ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3; }

// votes position p of read R.  out == nullptr: count only.  Returns the number of characters the position contributes;
// *present_out tells whether the pile holds anything here (run detection).
VOTE_FN int vote_position(const Ctx& c, const Read& R, uint32_t p, char* out, bool* present_out) {
  int ld0, maxq; bool filled;
  const bool present = scan_pos(c, R, p, &ld0, &maxq, &filled);
  *present_out = present;
  if (!present) return 0;
  const uint32_t w = c.P.w;
  int n = 0;
  if (filled) {                                  // uncorrected base in lower case
    if (out) { uint8_t b = (c.packed[R.boff + (p >> 2)] >> (6 - 2 * (p & 3))) & 3; out[0] = "acgt"[b]; }
    return 1;
  }
  int depth = ld0;
  if (ld0 == 0) {                                // only symbols inserted behind the end of a window: the depth is the one the right-to-left
    depth = -1;                                  // walk carries over from the nearest base column to the right of the same run (:2640-2700)
    for (uint32_t q = p + 1; q < R.span; ++q) {
      int l2, m2; bool f2;
      if (!scan_pos(c, R, q, &l2, &m2, &f2)) break;
      if (l2 > 0) { depth = l2; break; }
    }
  }
  const uint32_t lo = first_cover(c, R, p);
  for (int q = maxq; q >= 1; --q) {              // insertion columns, apre = -q, leftmost first
    int cnt[5] = {0, 0, 0, 0, 0}, ld = 0;
    for (uint32_t i = lo; i < R.we && c.win[i].astart <= p; ++i) {
      if (c.res[i].status != ST_OK) continue;
      const uint32_t e = c.ent[(uint64_t)i * (w + 1) + (p - c.win[i].astart)];
      if ((int)(e & 0xFF) < q) continue;
      cnt[sym_index(c.cons[(uint64_t)i * c.P.cons_stride + ((e >> 8) & 0x7F) - (uint32_t)q])]++; ++ld;
    }
    if (depth > ld) cnt[4] += depth - ld;
    const int ch = pick(cnt);
    if (ch) { if (out) out[n] = (char)ch; ++n; }
  }
  if (ld0 > 0) {                                 // base column
    int cnt[5] = {0, 0, 0, 0, 0};
    for (uint32_t i = lo; i < R.we && c.win[i].astart <= p; ++i) {
      if (c.res[i].status != ST_OK) continue;
      const uint32_t j = p - c.win[i].astart;
      if (j >= w) continue;
      const uint32_t e = c.ent[(uint64_t)i * (w + 1) + j];
      if (e & E_DEL) cnt[4]++; else cnt[sym_index(c.cons[(uint64_t)i * c.P.cons_stride + ((e >> 8) & 0x7F)])]++;
    }
    const int ch = pick(cnt);
    if (ch) { if (out) out[n] = (char)ch; ++n; }
  }
  return n;
}

}  // namespace dvote
