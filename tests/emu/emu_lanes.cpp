// TEST-ONLY build of the kernel source (daccord_b200/csrc/window_core.cuh) as a 32-lane host emulation
// (-DDCU_EMU -DDCU_EMU_LANES): each lane of the warp is a cooperative fiber (ucontext) with its own Ctx /
// WinState ("registers") on the shared workspace slab; fibers switch only inside warp collectives and
// wsync() (emu_xchg below), and the harness chooses the order in which lanes run between two such points:
//   schedule 0 ascending, 1 descending, >= 2 a fresh seeded shuffle at every scheduling pass.
// Any intra-warp data race of the lane-parallel code (a missing __syncwarp) makes the result depend on the
// schedule and shows up as a mismatch against the oracle; lanes that do not execute the same sequence of
// collectives (a __ballot_sync / __shfl_sync in divergent code: undefined on the GPU) are reported as a
// deadlock; lanes that disagree on the "uniform" per-window state (Result) are reported as well.
// Never linked into the product library.
#define DCU_EMU 1
#define DCU_EMU_LANES 1
#include "emu_builds.hpp"
#include <ucontext.h>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <map>
#include <dlfcn.h>

namespace {
constexpr int NL = 32;
int g_cur = -1;                       // lane that is running
constexpr size_t STACK_BYTES = 512 * 1024;
// context switch: on x86-64 a minimal callee-saved-register switch (swapcontext makes a sigprocmask system call per switch, which
// was half of the test's run time); elsewhere ucontext
#if defined(__x86_64__)
extern "C" void dcu_emu_swap(void** save_sp, void* new_sp);
asm(R"(
.text
.globl dcu_emu_swap
.type dcu_emu_swap,@function
dcu_emu_swap:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size dcu_emu_swap,.-dcu_emu_swap
)");
struct Fiber { void* sp; std::vector<char> stack; bool finished; };
void* g_sched_sp;
#define TO_SCHED(l) dcu_emu_swap(&g_fib[l].sp, g_sched_sp)
#define TO_LANE(l) dcu_emu_swap(&g_sched_sp, g_fib[l].sp)
#else
struct Fiber { ucontext_t uc; std::vector<char> stack; bool finished; };
ucontext_t g_sched;
#define TO_SCHED(l) swapcontext(&g_fib[l].uc, &g_sched)
#define TO_LANE(l) swapcontext(&g_sched, &g_fib[l].uc)
#endif
unsigned long long g_x[2][NL];        // double buffered exchange words: a lane can be at most one collective ahead of the slowest
unsigned g_gen = 0; int g_arrived = 0;
unsigned long long g_ncoll = 0;
void* g_last_site[NL];                // call site of every lane's last collective (printed when the lanes deadlock)

// what one lane runs
struct LaneJob { dcu::Ctx c; dcus::Ctx cs; dcu::Window W; dcu::Result r; uint8_t* cons; uint8_t* ops; };
LaneJob g_job[NL];
Fiber g_fib[NL];
bool g_smem_build = false;             // tier 2: the shared-memory build of the kernel source
void lane_main(int lane) {
  LaneJob& j = g_job[lane];
  if (g_smem_build) dcus::process_window(j.cs, j.W, j.r, j.cons, j.ops, lane);
  else dcu::process_window(j.c, j.W, j.r, j.cons, j.ops, lane);
  g_fib[lane].finished = true;
  TO_SCHED(lane);
}
void lane_entry() { lane_main(g_cur); abort(); }       // a finished lane is never resumed
std::map<void*, unsigned long long>* g_prof = nullptr;
uint64_t g_rng = 1;
inline uint32_t rnd() { g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(g_rng >> 33); }
}  // namespace

namespace dcub {
int emu_skip_sync_line = -1;
__attribute__((noinline)) const unsigned long long* emu_xchg(unsigned long long v) {
  const int lane = g_cur; const unsigned g = g_gen;
  if (g_prof && lane == 0) ++(*g_prof)[__builtin_return_address(0)];      // DCU_EMU_PROFILE: collectives per call site
  g_x[g & 1][lane] = v; g_last_site[lane] = __builtin_return_address(0);
  if (++g_arrived == NL) { g_arrived = 0; ++g_gen; ++g_ncoll; }
  else while (g_gen == g) TO_SCHED(lane);
  return g_x[g & 1];
}
}  // namespace dcub

// returns 0, or 1 = deadlock (lanes diverged around a collective), 2 = lanes disagree on the result record
static int run_warp(int schedule) {
  for (int l = 0; l < NL; ++l) {
    Fiber& f = g_fib[l];
    if (f.stack.empty()) f.stack.resize(STACK_BYTES);
    f.finished = false;
#if defined(__x86_64__)
    uintptr_t top = ((uintptr_t)f.stack.data() + f.stack.size()) & ~(uintptr_t)15;
    void** a = (void**)(top - 16);                   // return-address slot (16-byte aligned, so the entry sees rsp % 16 == 8 like after a call)
    a[0] = (void*)lane_entry;
    for (int q = 1; q <= 6; ++q) a[-q] = nullptr;    // the six callee-saved registers the switch pops
    f.sp = (void*)(a - 6);
#else
    getcontext(&f.uc);
    f.uc.uc_stack.ss_sp = f.stack.data(); f.uc.uc_stack.ss_size = f.stack.size(); f.uc.uc_link = &g_sched;
    makecontext(&f.uc, (void (*)())lane_main, 1, l);
#endif
  }
  g_arrived = 0;
  int order[NL];
  for (int l = 0; l < NL; ++l) order[l] = schedule == 1 ? NL - 1 - l : l;
  for (;;) {
    if (schedule >= 2) for (int i = NL - 1; i > 0; --i) std::swap(order[i], order[rnd() % (uint32_t)(i + 1)]);
    const unsigned gen0 = g_gen; const int arr0 = g_arrived; int nfin = 0, nfin0 = 0;
    for (int l = 0; l < NL; ++l) nfin0 += g_fib[l].finished;
    for (int q = 0; q < NL; ++q) {
      const int l = order[q];
      if (g_fib[l].finished) continue;
      g_cur = l;
      TO_LANE(l);
    }
    for (int l = 0; l < NL; ++l) nfin += g_fib[l].finished;
    if (nfin == NL) break;
    if (g_gen == gen0 && g_arrived == arr0 && nfin == nfin0) return 1;      // a whole pass without progress
    if (nfin > 0 && g_arrived + nfin == NL && g_gen == gen0) return 1;      // the rest waits for lanes that are gone
  }
  for (int l = 1; l < NL; ++l) if (memcmp(&g_job[l].r, &g_job[0].r, sizeof(dcu::Result)) != 0) return 2;
  return 0;
}

extern "C" int emu_lanes_run_batch(const dcu_params* prm, const uint8_t* packed, const dcu_window* win, uint64_t nwin, const dcu_slice* sl,
                                   dcu_result* res, uint8_t* cons, uint8_t* ops, int tier, uint64_t* noverflow, int schedule, uint64_t seed,
                                   uint64_t* ncollectives) {
  dcu_host::HostTables HT;
  int maxS = 4, maxB = 64;
  for (uint64_t i = 0; i < nwin; ++i) {
    int b = 0;
    for (uint32_t j = 0; j < win[i].slice_cnt; ++j) b += sl[win[i].slice_begin + j].len;
    maxS = std::max<int>(maxS, win[i].slice_cnt); maxB = std::max(maxB, b);
  }
  dcu_host::build_tables((int)prm->w, prm->p_i, prm->p_d, prm->est_cor, (int)prm->k_lo, (int)prm->k_hi, maxS + 2, HT);
  dcu::Caps caps = emu::caps_for(tier, (int)prm->w, maxS, maxB);
  g_smem_build = tier == 2;
  dcu::Layout L; dcus::Layout LS; dcu::make_layout(caps, L); dcus::make_layout(caps, LS);
  std::vector<uint8_t> slab((g_smem_build ? LS.bytes : L.bytes) + 64), arena(LS.sbytes + 64);
  dcu::Tables T; dcu::Params P;
  emu::tables_for(HT, T); emu::params_for(prm, tier, P);
  if (g_smem_build) emu::BuildS::globals(LS, caps, T, P); else emu::BuildG::globals(L, caps, T, P);
  g_rng = seed * 2 + 1; g_ncoll = 0;
  { const char* e = getenv("DCU_EMU_SKIP_SYNC_LINE"); dcu::emu_skip_sync_line = e ? atoi(e) : -1; }
  if (getenv("DCU_EMU_PROFILE")) g_prof = new std::map<void*, unsigned long long>();
  uint64_t nov = 0;
  for (uint64_t i = 0; i < nwin; ++i) {
    memset(cons + i * DCU_CONS_STRIDE, 0, DCU_CONS_STRIDE); memset(ops + i * DCU_OPS_STRIDE, 0, DCU_OPS_STRIDE);
    for (int l = 0; l < NL; ++l) {
      LaneJob& j = g_job[l];
      memset(&j.c, 0, sizeof(j.c)); memset(&j.cs, 0, sizeof(j.cs));
      j.c.ws.base = slab.data(); j.c.vsq = T.VSq; j.c.vs_sm = 0; j.c.packed = packed; j.c.sl = (const dcu::Slice*)sl;
      j.cs.ws.base = slab.data(); j.cs.ws.sm = arena.data(); j.cs.vsq = T.VSq; j.cs.vs_sm = 0; j.cs.packed = packed; j.cs.sl = (const dcu::Slice*)sl;
      j.c.epoch = j.cs.epoch = (unsigned long long)(i + 1) << 32;      // tags of the forward slot records: unique per window, as the kernel's per-warp counter is
      memcpy(&j.W, &win[i], sizeof(j.W));
      memset(&j.r, 0, sizeof(j.r));
      j.cons = cons + i * DCU_CONS_STRIDE; j.ops = ops + i * DCU_OPS_STRIDE;
    }
    int rc = run_warp(schedule);
    if (rc) {
      fprintf(stderr, "emu_lanes: window %lu: %s\n", (unsigned long)i, rc == 1 ? "deadlock: lanes diverged around a warp collective" : "lanes disagree on the result record");
      if (rc == 1 && getenv("DCU_EMU_DEBUG")) { Dl_info di; dladdr((void*)&emu_lanes_run_batch, &di); for (int l = 0; l < NL; ++l) fprintf(stderr, "  lane %d finished %d last collective at +0x%lx\n", l, (int)g_fib[l].finished, (unsigned long)((char*)g_last_site[l] - (char*)di.dli_fbase)); }
      return 100 + rc;
    }
    memcpy(&res[i], &g_job[0].r, sizeof(dcu::Result));
    if (g_job[0].r.status == dcu::ST_OVERFLOW) ++nov;
  }
  if (noverflow) *noverflow = nov;
  if (ncollectives) *ncollectives = g_ncoll;
  if (g_prof) {      // "offset count" lines for addr2line -f -i -e libemu_lanes.so (tools/lane_collectives.py)
    FILE* fp = fopen(getenv("DCU_EMU_PROFILE"), "w");
    Dl_info di; dladdr((void*)&emu_lanes_run_batch, &di);
    if (fp) { for (auto& kv : *g_prof) fprintf(fp, "%lx %llu\n", (unsigned long)((char*)kv.first - (char*)di.dli_fbase), kv.second); fclose(fp); }
    delete g_prof; g_prof = nullptr;
  }
  return 0;
}
