// TEST-ONLY build of the kernel source (daccord_b200/csrc/window_core.cuh) as a single-lane host
// emulation (-DDCU_EMU): lets the parity tests exercise the product's per-window logic against the
// oracle in a container without a GPU.  Never linked into the product library.
#define DCU_EMU 1
#include "emu_builds.hpp"
#include "../../include/daccord_b200.h"
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <chrono>

#ifdef DCU_EMU_TRACE
extern "C" void trace_begin_window(const uint8_t* base, uint64_t bytes, const uint32_t* off, int nfields);      // tests/emu/trace_rt.cpp
extern "C" void trace_end_window();
#endif
template <class B> static int run_batch(const dcu_params* prm, const uint8_t* packed, const dcu_window* win, uint64_t nwin, const dcu_slice* sl,
                                        dcu_result* res, uint8_t* cons, uint8_t* ops, int tier, uint64_t* noverflow) {
  dcu_host::HostTables HT;
  int maxS = 4, maxB = 64;
  for (uint64_t i = 0; i < nwin; ++i) {
    int b = 0;
    for (uint32_t j = 0; j < win[i].slice_cnt; ++j) b += sl[win[i].slice_begin + j].len;
    maxS = std::max<int>(maxS, win[i].slice_cnt); maxB = std::max(maxB, b);
  }
  dcu_host::build_tables((int)prm->w, prm->p_i, prm->p_d, prm->est_cor, (int)prm->k_lo, (int)prm->k_hi, maxS + 2, HT);
  dcu::Caps caps = emu::caps_for(tier, (int)prm->w, maxS, maxB);
  typename B::Layout L; B::layout(caps, L);
  std::vector<uint8_t> slab(L.bytes + 64), arena(L.sbytes + 64);
  dcu::Tables T; dcu::Params P;
  emu::tables_for(HT, T); emu::params_for(prm, tier, P);
  B::globals(L, caps, T, P);
  typename B::Ctx c;
  B::bind(c, slab.data(), arena.data()); c.vsq = T.VSq; c.vs_sm = 0; c.epoch = 0;      // (the slab starts zeroed: no forward slot record carries a tag yet)
  c.packed = packed; c.sl = (const dcu::Slice*)sl;
  uint64_t nov = 0;
  for (uint64_t i = 0; i < nwin; ++i) {
    dcu::Result r;
    dcu::Window W; memcpy(&W, &win[i], sizeof(W));
    memset(cons + i * DCU_CONS_STRIDE, 0, DCU_CONS_STRIDE); memset(ops + i * DCU_OPS_STRIDE, 0, DCU_OPS_STRIDE);
#ifdef DCU_EMU_STATS
    for (int q = 0; q < 16; ++q) { g_peak[q] = 0; dcu::g_phase_ns[q] = 0; dcu::g_phase_calls[q] = 0; }
    const auto t0 = std::chrono::steady_clock::now();
#endif
#ifdef DCU_EMU_TRACE
    trace_begin_window(slab.data(), L.bytes, L.off, (int)dcu::F_COUNT);
#endif
    B::process(c, W, r, cons + i * DCU_CONS_STRIDE, ops + i * DCU_OPS_STRIDE, 0);
#ifdef DCU_EMU_TRACE
    trace_end_window();
#endif
#ifdef DCU_EMU_STATS
    g_peak[9] = (long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();    // single-lane emulation time of the window
    if (const char* fn = getenv("DCU_FOOTPRINT_OUT")) {      // per-window peaks of the workspace counters (tools/footprint.py)
      static FILE* fp = nullptr;
      if (!fp) fp = fopen(fn, "w");
      if (fp) { for (int q = 0; q < 15; ++q) fprintf(fp, "%ld ", g_peak[q]); for (int q = 0; q < 11; ++q) fprintf(fp, "%ld %ld%c", dcu::g_phase_ns[q], dcu::g_phase_calls[q], q == 10 ? '\n' : ' '); fflush(fp); }
    }
#endif
    memcpy(&res[i], &r, sizeof(r));
    if (r.status == dcu::ST_OVERFLOW) ++nov;
  }
#ifdef DCU_EMU_STATS
  fprintf(stderr, "windows %lu traverse calls %ld pairs %ld narp/pair %.1f nds/pair %.1f nn %.1f nrl %.1f links/pair %.1f\n", (unsigned long)nwin, dcu::g_stats[6], dcu::g_stats[0], (double)dcu::g_stats[1] / dcu::g_stats[0], (double)dcu::g_stats[2] / dcu::g_stats[0], (double)dcu::g_stats[3] / dcu::g_stats[0], (double)dcu::g_stats[4] / dcu::g_stats[0], (double)dcu::g_stats[5] / dcu::g_stats[0]);
  fprintf(stderr, "pairs without an accepted reverse path %ld, pairs that left the candidate heap untouched %ld; last k-mer unreachable from the first %ld (dead and unreachable %ld, dead and unreachable or beyond lmax %ld, alive though unreachable / too far %ld)\n", dcu::g_stats[7], dcu::g_stats[8], dcu::g_stats[9], dcu::g_stats[10], dcu::g_stats[11], dcu::g_stats[12]);
  for (int i = 0; i < 16; ++i) dcu::g_stats[i] = 0;
#endif
  if (noverflow) *noverflow = nov;
  return 0;
}
// tier 0 / 1: HBM build with the capacities of the library's first / second overflow pass; tier 2: shared-memory build
extern "C" int emu_run_batch(const dcu_params* prm, const uint8_t* packed, const dcu_window* win, uint64_t nwin, const dcu_slice* sl,
                             dcu_result* res, uint8_t* cons, uint8_t* ops, int tier, uint64_t* noverflow) {
  return tier == 3 ? run_batch<emu::BuildH>(prm, packed, win, nwin, sl, res, cons, ops, tier, noverflow)
       : tier == 2 ? run_batch<emu::BuildS>(prm, packed, win, nwin, sl, res, cons, ops, tier, noverflow)
                   : run_batch<emu::BuildG>(prm, packed, win, nwin, sl, res, cons, ops, tier, noverflow);
}
// product table builder exposed for the table-parity test
extern "C" int64_t emu_get_tables(const dcu_params* prm, int which, double* out, int64_t cap, int klimn) {
  dcu_host::HostTables HT;
  dcu_host::build_tables((int)prm->w, prm->p_i, prm->p_d, prm->est_cor, (int)prm->k_lo, (int)prm->k_hi, klimn, HT);
  std::vector<double> v;
  if (which == 0) v = HT.DPn; else if (which == 1) v = HT.DPsq;
  else if (which == 2) { for (int l = 0; l < HT.NP; ++l) for (int q = 0; q < HT.MS; ++q) v.push_back((double)HT.VSq[(size_t)q * HT.NP + l]); }
  else if (which == 3) for (int i = 0; i < HT.MS; ++i) { v.push_back(HT.suplo[i]); v.push_back(HT.suphi[i]); }
  else if (which == 4) for (auto x : HT.klim) v.push_back((double)x);
  else if (which == 5) { v.push_back(HT.NP); v.push_back(HT.MS); }
  if ((int64_t)v.size() <= cap) memcpy(out, v.data(), v.size() * sizeof(double));
  return (int64_t)v.size();
}

// ---- host emulation of the GPU piling stage (pile_core.cuh), for the parity test against the host piler
#include "../../daccord_b200/csrc/pile_host.hpp"
extern "C" int emu_pile(const dcu_overlap* ovl, uint64_t novl, const uint16_t* trace, uint64_t ntrace, int32_t tspace, const uint8_t* packed,
                        const uint64_t* read_boff, const uint32_t* read_len, uint64_t nreads, uint32_t w, uint32_t a, uint64_t maxalign,
                        dcu_window* win_out, uint64_t win_cap, dcu_slice* sl_out, uint64_t sl_cap, uint64_t* nwin, uint64_t* nsl) {
  dpile::Prep P;
  if (!dpile::prepare(ovl, novl, ntrace, tspace, w, a, nreads, read_len, P)) { fprintf(stderr, "emu_pile: %s\n", P.err.c_str()); return 1; }
  dpile::Params prm; prm.tspace = tspace; prm.w = w; prm.a = a; prm.maxalign = maxalign;
  std::vector<uint32_t> tile_b(P.ntiles + 1), bm(P.nbm + 1, 0xDEADBEEFu);
  for (auto& o : P.ovl) dpile::pile_tile_starts(o, trace, tile_b.data());
  std::vector<dpile::U128> PV(dpile::PILE_MAXB + 1), MV(dpile::PILE_MAXB + 1), PH(dpile::PILE_MAXB + 1), MH(dpile::PILE_MAXB + 1);
  for (size_t r = 0; r < P.reads.size(); ++r) {
    const uint32_t l = P.reads[r].maxaepos, s0 = l >= w ? l - w : 0, s1 = l;
    for (uint64_t i = P.reads[r].ovl_begin; i < P.reads[r].ovl_end; ++i)
      for (int t = 0; t < P.ovl[i].ntiles; ++t) dpile::pile_align_tile(P.ovl[i], t, prm, trace, tile_b.data(), packed, read_boff, read_len, s0, s1, bm.data(), PV.data(), MV.data(), PH.data(), MH.data());
  }
  std::vector<unsigned long long> keys(novl + 1);
  for (size_t r = 0; r < P.reads.size(); ++r) dpile::pile_order(P.reads[r], P.ovl.data(), P.minerate[r], P.ediv[r], keys.data() + P.reads[r].ovl_begin);
  // pass 0: slices per candidate window; exclusive scans = what the block scans of the kernels produce
  std::vector<uint32_t> cnt(P.ncand + 1, 0);
  for (size_t r = 0; r < P.reads.size(); ++r)
    for (uint32_t y = 0; y < P.reads[r].nwin; ++y)
      cnt[P.reads[r].win_off + y] = (uint32_t)dpile::pile_window(P.reads[r], y, P.ovl.data(), prm, bm.data(), read_boff, read_len, keys.data() + P.reads[r].ovl_begin, P.read_id[r], nullptr, nullptr, 0);
  uint64_t tw = 0, ts = 0;
  for (uint64_t i = 0; i < P.ncand; ++i) { tw += cnt[i] ? 1 : 0; ts += cnt[i]; }
  *nwin = tw; *nsl = ts;
  if (tw > win_cap || ts > sl_cap) return 3;
  uint64_t wo = 0, so = 0;
  for (size_t r = 0; r < P.reads.size(); ++r)
    for (uint32_t y = 0; y < P.reads[r].nwin; ++y) {
      const uint32_t n = cnt[P.reads[r].win_off + y];
      if (!n) continue;
      int rc = dpile::pile_window(P.reads[r], y, P.ovl.data(), prm, bm.data(), read_boff, read_len, keys.data() + P.reads[r].ovl_begin, P.read_id[r],
                                  (dpile::Win*)win_out + wo, (dpile::Sl*)sl_out + so, (uint32_t)so);
      if (rc != (int)n) return 2;
      ++wo; so += n;
    }
  return 0;
}

// ---- host emulation of the GPU pile vote (vote_core.cuh), for the parity test against host/vote.hpp
#include "../../daccord_b200/csrc/vote_host.hpp"
extern "C" int emu_vote(const dcu_window* win, const dcu_result* res, const uint8_t* cons, const uint8_t* ops, uint64_t nwin, uint32_t w, int producefull, uint64_t minlen,
                        const uint8_t* packed, const uint64_t* read_boff, const uint32_t* read_len, uint64_t nreads,
                        dcu_segment* seg_out, uint64_t seg_cap, char* chars_out, uint64_t chars_cap, uint64_t* nseg, uint64_t* nchars) {
  dvote::Layout L;
  if (!dvote::layout_reads(win, nwin, w, producefull != 0, read_boff, read_len, nreads, L)) { fprintf(stderr, "emu_vote: %s\n", L.err.c_str()); return 1; }
  dvote::Params vp; vp.w = w; vp.cons_stride = DCU_CONS_STRIDE; vp.ops_stride = DCU_OPS_STRIDE; vp.producefull = producefull ? 1 : 0;
  std::vector<uint16_t> ent(nwin * (w + 1) + 1, 0xFFFF);
  for (uint64_t i = 0; i < nwin; ++i)
    if (res[i].status == dvote::ST_OK && !dvote::vote_window_table(((const dvote::Res*)res)[i], ops + i * DCU_OPS_STRIDE, w, ent.data() + i * (w + 1))) return 2;
  dvote::Ctx c; c.win = (const dvote::Win*)win; c.res = (const dvote::Res*)res; c.cons = cons; c.ent = ent.data(); c.packed = packed; c.P = vp;
  std::vector<uint8_t> flag(L.npos + 1, 0); std::vector<uint64_t> off(L.npos + 1, 0);
  uint64_t total = 0;
  for (size_t r = 0; r < L.reads.size(); ++r)
    for (uint32_t p = 0; p < L.reads[r].span; ++p) {
      bool pr; int n = dvote::vote_position(c, L.reads[r], p, nullptr, &pr);
      if (n > 127) return 3;
      const uint64_t idx = L.reads[r].pos_off + p;
      flag[idx] = (uint8_t)(n | (pr ? 0x80 : 0)); off[idx] = total; total += (uint64_t)n;
    }
  *nchars = total;
  if (total > chars_cap) return 4;
  std::vector<dvote::Bound> B;
  for (size_t r = 0; r < L.reads.size(); ++r)
    for (uint32_t p = 0; p < L.reads[r].span; ++p) {
      const uint64_t idx = L.reads[r].pos_off + p;
      if (!(flag[idx] & 0x80)) continue;
      const uint32_t n = flag[idx] & 0x7F; bool pr;
      if (n) dvote::vote_position(c, L.reads[r], p, chars_out + off[idx], &pr);
      const bool left = p > 0 && (flag[idx - 1] & 0x80), right = p + 1 < L.reads[r].span && (flag[idx + 1] & 0x80);
      if (!left) B.push_back(dvote::Bound{off[idx], (uint32_t)r, p, 0u, 0u});
      if (!right) B.push_back(dvote::Bound{off[idx] + n, (uint32_t)r, p, 1u, 0u});
    }
  std::vector<dcu_segment> seg; std::string err;
  if (!dvote::pair_bounds(B, L, producefull != 0, minlen, seg, err)) { fprintf(stderr, "emu_vote: %s\n", err.c_str()); return 5; }
  *nseg = seg.size();
  if (seg.size() > seg_cap) return 6;
  if (!seg.empty()) memcpy(seg_out, seg.data(), seg.size() * sizeof(dcu_segment));
  return 0;
}
