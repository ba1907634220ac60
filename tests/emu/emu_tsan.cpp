// TEST-ONLY: the kernel source (daccord_b200/csrc/window_core.cuh, -DDCU_EMU -DDCU_EMU_LANES) with the 32 lanes of a warp as 32 OS
// threads under ThreadSanitizer.  Warp collectives and wsync() are pthread barriers, atomicCAS / atomicAdd are real atomics, so every
// other pair of conflicting accesses by two lanes between two barriers is reported by TSan with both source lines -- the CPU stand-in for
// a race checker on the per-warp workspace (which lives in global memory, where compute-sanitizer's racecheck does not look).
// Stand-alone executable (the TSan runtime cannot be loaded into the Python process): reads a batch file written by tests/common.py
// (write_batch_file), runs every window, writes the result records / consensus / placement ops.  Never linked into the product.
//   g++ -fsanitize=thread -O1 -g -std=c++17 -ffp-contract=off -pthread -o emu_tsan emu_tsan.cpp ;  ./emu_tsan batch.bin out.bin tier
#define DCU_EMU 1
#define DCU_EMU_LANES 1
#include "emu_builds.hpp"
#include <pthread.h>
#include <thread>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace {
constexpr int NL = 32;
pthread_barrier_t g_bar;
unsigned long long g_x[2][NL];
thread_local int t_lane = 0;
thread_local unsigned t_gen = 0;
}
namespace dcub {
int emu_skip_sync_line = -1;
const unsigned long long* emu_xchg(unsigned long long v) {
  const unsigned g = t_gen++;
  g_x[g & 1][t_lane] = v;               // double buffered: writing this buffer again needs two more barriers, by then every lane has read it
  pthread_barrier_wait(&g_bar);
  return g_x[g & 1];
}
}

template <class T> static bool rd(FILE* f, std::vector<T>& v, uint64_t n) { v.resize(n); return n == 0 || fread(v.data(), sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: emu_tsan batch.bin out.bin tier\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  dcu_params prm; uint64_t hdr[3];
  if (fread(&prm, sizeof(prm), 1, f) != 1 || fread(hdr, 8, 3, f) != 3) return 2;
  std::vector<uint8_t> packed; std::vector<dcu_window> win; std::vector<dcu_slice> sl;
  if (!rd(f, packed, hdr[0]) || !rd(f, win, hdr[1]) || !rd(f, sl, hdr[2])) return 2;
  fclose(f);
  const int tier = atoi(argv[3]);
  { const char* e = getenv("DCU_EMU_SKIP_SYNC_LINE"); dcu::emu_skip_sync_line = e ? atoi(e) : -1; }      // mutation testing (tools/lane_mutants.py --tsan)
  const uint64_t nwin = win.size();
  dcu_host::HostTables HT;
  int maxS = 4, maxB = 64;
  for (uint64_t i = 0; i < nwin; ++i) {
    int b = 0;
    for (uint32_t j = 0; j < win[i].slice_cnt; ++j) b += sl[win[i].slice_begin + j].len;
    maxS = std::max<int>(maxS, win[i].slice_cnt); maxB = std::max(maxB, b);
  }
  dcu_host::build_tables((int)prm.w, prm.p_i, prm.p_d, prm.est_cor, (int)prm.k_lo, (int)prm.k_hi, maxS + 2, HT);
  dcu::Caps caps = emu::caps_for(tier, (int)prm.w, maxS, maxB);
  const bool smem_build = tier == 2;                  // the shared-memory build of the kernel source (its arena is a host buffer here)
  dcu::Layout L; dcus::Layout LS; dcu::make_layout(caps, L); dcus::make_layout(caps, LS);
  std::vector<uint8_t> slab((smem_build ? LS.bytes : L.bytes) + 64), arena(LS.sbytes + 64);
  dcu::Tables T; dcu::Params P;
  emu::tables_for(HT, T); emu::params_for(&prm, tier, P); P.defer_ff = 0;
  if (smem_build) emu::BuildS::globals(LS, caps, T, P); else emu::BuildG::globals(L, caps, T, P);
  std::vector<dcu_result> res(nwin); std::vector<uint8_t> cons(nwin * DCU_CONS_STRIDE, 0), ops(nwin * DCU_OPS_STRIDE, 0);
  std::vector<dcu::Result> lane_res((size_t)NL * nwin);
  pthread_barrier_init(&g_bar, nullptr, NL);
  std::vector<std::thread> th;
  for (int l = 0; l < NL; ++l) th.emplace_back([&, l]() {
    t_lane = l; t_gen = 0;
    for (uint64_t i = 0; i < nwin; ++i) {
      dcu::Window W; memcpy(&W, &win[i], sizeof(W));
      dcu::Result r; memset(&r, 0, sizeof(r));
      if (smem_build) {
        dcus::Ctx c; memset(&c, 0, sizeof(c));
        c.ws.base = slab.data(); c.ws.sm = arena.data(); c.vsq = T.VSq; c.vs_sm = 0; c.epoch = (unsigned long long)(i + 1) << 32; c.packed = packed.data(); c.sl = (const dcu::Slice*)sl.data();
        dcus::process_window(c, W, r, cons.data() + i * DCU_CONS_STRIDE, ops.data() + i * DCU_OPS_STRIDE, l);
      } else {
        dcu::Ctx c; memset(&c, 0, sizeof(c));
        c.ws.base = slab.data(); c.vsq = T.VSq; c.vs_sm = 0; c.epoch = (unsigned long long)(i + 1) << 32; c.packed = packed.data(); c.sl = (const dcu::Slice*)sl.data();
        dcu::process_window(c, W, r, cons.data() + i * DCU_CONS_STRIDE, ops.data() + i * DCU_OPS_STRIDE, l);
      }
      lane_res[(size_t)l * nwin + i] = r;
      dcu::emu_xchg(0);                 // the kernel's __syncwarp() after a window is published
    }
  });
  for (auto& t : th) t.join();
  int disagree = 0;
  for (uint64_t i = 0; i < nwin; ++i) {
    memcpy(&res[i], &lane_res[i], sizeof(dcu_result));
    for (int l = 1; l < NL; ++l) if (memcmp(&lane_res[(size_t)l * nwin + i], &lane_res[i], sizeof(dcu::Result)) != 0) ++disagree;
  }
  f = fopen(argv[2], "wb");
  if (!f) return 2;
  fwrite(res.data(), sizeof(dcu_result), nwin, f); fwrite(cons.data(), 1, cons.size(), f); fwrite(ops.data(), 1, ops.size(), f);
  fclose(f);
  if (disagree) { fprintf(stderr, "emu_tsan: %d lane result records differ from lane 0\n", disagree); return 3; }
  return 0;
}
