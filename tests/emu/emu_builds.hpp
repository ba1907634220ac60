// TEST-ONLY: both builds of the kernel source for the host emulations.  tier 0 / 1 = HBM build (namespace dcu, first / second
// overflow pass of the library), tier 2 = shared-memory build (namespace dcus; its arena is a host buffer here), tier 3 = hybrid build
// (namespace dcuh: only the k-mer table in the arena).
#pragma once
#include "../../daccord_b200/csrc/window_core.cuh"
#define DCU_NS dcus
#define DCU_TIER_SMEM 1
#include "../../daccord_b200/csrc/window_core.cuh"
#define DCU_NS dcuh
#define DCU_TIER_SMEM 2
#include "../../daccord_b200/csrc/window_core.cuh"
#include "../../daccord_b200/csrc/host_tables.hpp"
#include "../../daccord_b200/csrc/host_caps.hpp"
#include "../../include/daccord_b200.h"
#include <vector>
#include <cstdlib>

namespace emu {
struct BuildG {
  typedef dcu::Ctx Ctx; typedef dcu::Layout Layout; typedef dcu::WinState WinState;
  static void layout(const dcu::Caps& c, Layout& L) { dcu::make_layout(c, L); }
  static void globals(const Layout& L, const dcu::Caps& c, const dcu::Tables& T, const dcu::Params& P) { dcu::g_layout = L; dcu::g_cap = c; dcu::g_T = T; dcu::g_P = P; }
  static void bind(Ctx& c, uint8_t* slab, uint8_t*) { c.ws.base = slab; }
  static void process(Ctx& c, const dcu::Window& W, dcu::Result& r, uint8_t* cons, uint8_t* ops, int lane) { dcu::process_window(c, W, r, cons, ops, lane); }
};
struct BuildS {
  typedef dcus::Ctx Ctx; typedef dcus::Layout Layout; typedef dcus::WinState WinState;
  static void layout(const dcu::Caps& c, Layout& L) { dcus::make_layout(c, L); }
  static void globals(const Layout& L, const dcu::Caps& c, const dcu::Tables& T, const dcu::Params& P) { dcus::g_layout = L; dcus::g_cap = c; dcus::g_T = T; dcus::g_P = P; }
  static void bind(Ctx& c, uint8_t* slab, uint8_t* arena) { c.ws.base = slab; c.ws.sm = arena; }
  static void process(Ctx& c, const dcu::Window& W, dcu::Result& r, uint8_t* cons, uint8_t* ops, int lane) { dcus::process_window(c, W, r, cons, ops, lane); }
};
struct BuildH {
  typedef dcuh::Ctx Ctx; typedef dcuh::Layout Layout; typedef dcuh::WinState WinState;
  static void layout(const dcu::Caps& c, Layout& L) { dcuh::make_layout(c, L); }
  static void globals(const Layout& L, const dcu::Caps& c, const dcu::Tables& T, const dcu::Params& P) { dcuh::g_layout = L; dcuh::g_cap = c; dcuh::g_T = T; dcuh::g_P = P; }
  static void bind(Ctx& c, uint8_t* slab, uint8_t* arena) { c.ws.base = slab; c.ws.sm = arena; }
  static void process(Ctx& c, const dcu::Window& W, dcu::Result& r, uint8_t* cons, uint8_t* ops, int lane) { dcuh::process_window(c, W, r, cons, ops, lane); }
};
inline dcu::Caps caps_for(int tier, int w, int maxS, int maxB) { return tier == 3 ? dcu_host::make_caps_hybrid(w, maxS, maxB) : (tier == 2 ? dcu_host::make_caps_smem(w, maxS, maxB) : dcu_host::make_caps(tier, w, maxS, maxB)); }
inline void params_for(const dcu_params* prm, int tier, dcu::Params& P) {
  P.w = (int)prm->w; P.k_lo = (int)prm->k_lo; P.k_hi = (int)prm->k_hi; P.minff = prm->min_ff; P.maxff = prm->max_ff;
  P.mincov = (int)prm->min_cov; P.check = prm->est_cor != 0.0; P.eminrate = prm->max_err;
  P.defer_ff = (tier == 0 && getenv("DCU_DEFER_FF")) ? 1 : 0;
  { const char* e = getenv("DCU_POSCACHE"); P.poscache = e ? atoi(e) : 1; }        // same experimental switch as the library's first pass
}
inline void tables_for(const dcu_host::HostTables& HT, dcu::Tables& T) {
  T.DPn = HT.DPn.data(); T.DPsq = HT.DPsq.data(); T.VSq = HT.VSq.data(); T.suplo = HT.suplo.data(); T.suphi = HT.suphi.data();
  T.klim = HT.klim.data(); T.NP = HT.NP; T.MS = HT.MS; T.KLIMN = HT.KLIMN;
}
}  // namespace emu
