// host_caps.hpp -- workspace capacities of the two kernel tiers (see DESIGN.md "workspace tiers").
// tier 0: sized for the common case (filterfreq 2, a handful of unitigs) so that a warp's slab stays
//         cache resident; tier 1: sized for anything the 16-bit indices of the kernel can address.
// A window that overflows tier 0 is re-run in tier 1; one that overflows tier 1 is reported as an error.
#pragma once
#include "window_types.cuh"
#include <cstdlib>
namespace dcu_host {
inline int ceil_pow2_log(int v) { int l = 4; while ((1 << l) < v) ++l; return l; }
inline dcu::Caps make_caps(int tier, int w, int maxS, int maxB) {
  dcu::Caps c;
  c.S = maxS < 4 ? 4 : maxS;
  c.B = maxB < 64 ? 64 : maxB;
  c.NI = c.B;
  c.EX = tier == 0 ? 1024 : 4096;                  // gap filler extras (the rare windows with more than 256 used to cost a whole extra launch)
  c.LOGH = ceil_pow2_log((c.NI + (tier == 0 ? 256 : c.EX)) * 3 / 2);
  c.H = 1 << c.LOGH;
  c.BL = w + 8;
  c.HEAVY = 0;
  if (tier == 0 && getenv("DCU_T0_SMALL")) {
    // measurement knob (tools/round2_first_run.sh): first-pass capacities near the p99.9 of the 40x bench workload instead of 5x above it.
    // The slab of a warp then spans ~1/4 of the address range (fewer 2 MB pages live per SM), at the price of more windows for the second pass.
    c.NN = 1024; c.ST = 512; c.SL = 2048; c.SF = 6144; c.RL = 512; c.RP = 512; c.FP = 512; c.SI = 512; c.KW = 2; c.HEAVY = 0;
  } else if (tier == 0) {
    c.NN = 4096; c.ST = 2048; c.SL = 16384; c.SF = 32768; c.RL = 4096; c.RP = 4096; c.FP = 4096; c.SI = 4096; c.KW = 2; c.HEAVY = 0;       // HEAVY > 0 hands graphs with more nodes to the free-running pass (measured: no gain, profiles/r01_summary.md)
  } else {
    c.NN = c.NI + c.EX; if (c.NN > 65000) c.NN = 65000;
    c.ST = 8192; c.SL = 65000; c.SF = 262144; c.RL = 32768; c.RP = 32768; c.FP = 32768; c.SI = 32768; c.KW = 2;
  }
  if (c.NN > c.NI + c.EX) c.NN = c.NI + c.EX;
  c.STP = 1 << ceil_pow2_log(c.ST);
  c.RLP = 1 << ceil_pow2_log(c.RL);
  c.BW = c.B / 16 + c.S + 1;                       // every slice starts on a word
  if (c.SL < c.NN) c.SL = c.NN;                    // the per-node fill counters of build_nodes alias the link array
  c.NBITS = 0; c.LOGNB = 0;                        // no pre-filter in the HBM build: its table takes every k-mer
  return c;
}
// capacities of the shared-memory build (dcus): sized for the common window of the batch, so that as many warps as possible
// share an SM; what does not fit is handed to the HBM passes (tier 0, then tier 1).  The hot fields (marked S in
// window_core.cuh) are laid out in the warp's shared-memory arena from these numbers.  DCU_S_* environment variables override
// single values (measurement knobs, results do not depend on them).
inline int env_or(const char* name, int v) { const char* e = getenv(name); return (e && atoi(e) > 0) ? atoi(e) : v; }
inline dcu::Caps make_caps_smem(int w, int maxS, int maxB) {
  dcu::Caps c;
  c.S = maxS < 4 ? 4 : (maxS > 80 ? 80 : maxS);
  c.S = env_or("DCU_S_S", c.S);
  int b = maxB < 64 ? 64 : (maxB > 3072 ? 3072 : maxB);
  c.BW = env_or("DCU_S_BW", b / 16 + c.S / 2 + 1);
  c.B = 16 * c.BW;
  c.NI = env_or("DCU_S_NI", b * 3 / 5 < 256 ? 256 : b * 3 / 5);          // instances of the nodes kept by the filter (40x: median 881, p99 1387 of ~1900 bases)
  c.EX = 64;
  c.LOGH = env_or("DCU_S_LOGH", b >= 1024 ? 10 : 9);
  c.H = 1 << c.LOGH;
  c.LOGNB = env_or("DCU_S_LOGNB", b >= 1024 ? 13 : 12);
  c.NBITS = 1 << c.LOGNB;
  c.NN = env_or("DCU_S_NN", b >= 1024 ? 384 : 256);
  c.SL = env_or("DCU_S_SL", c.NN + c.NN / 2);
  c.ST = 256; c.SF = 4096; c.RL = 512; c.RP = 512; c.FP = 512; c.SI = 512; c.KW = 2; c.HEAVY = 0;
  c.BL = w + 8;
  if (c.SL < c.NN) c.SL = c.NN;
  c.STP = 1 << ceil_pow2_log(c.ST);
  c.RLP = 1 << ceil_pow2_log(c.RL);
  return c;
}
// capacities of the hybrid build (dcuh): tier-0 capacities in the HBM slab, but the k-mer table in shared memory -- 512 slots behind the
// two-bitmap pre-filter (it then holds the k-mers seen twice: ~250 of a 40x window), 5 KB per warp, so that all 32 warps of an SM keep
// theirs.  A window whose table fills up, or that needs the filter frequency 1 graph, is handed to the plain HBM pass.
inline dcu::Caps make_caps_hybrid(int w, int maxS, int maxB) {
  dcu::Caps c = make_caps(0, w, maxS, maxB);
  c.EX = 64;
  c.LOGH = env_or("DCU_H_LOGH", 9); c.H = 1 << c.LOGH;
  c.LOGNB = env_or("DCU_H_LOGNB", 13); c.NBITS = 1 << c.LOGNB;
  if (c.NN > c.NI + c.EX) c.NN = c.NI + c.EX;
  if (c.NN > 32000) c.NN = 32000;                  // node ids share 16-bit slot values with the counts
  if (c.SL < c.NN) c.SL = c.NN;
  return c;
}
}  // namespace dcu_host
