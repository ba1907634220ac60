// TEST / PLANNING ONLY: a stand-in for the ThreadSanitizer runtime.  tests/emu/emu.cpp is compiled with -fsanitize=thread (which makes the
// compiler call __tsan_readN / __tsan_writeN before every memory access) but linked against THIS file instead of libtsan: every access that
// falls into the current window's workspace slab is attributed to the slab field it hits (window_core.cuh DCU_WS_FIELDS).  Per window the
// distinct 32-byte sectors per field are counted -- the bytes a window really needs from the memory system, by field -- next to load / store
// counts.  Used by tools/field_traffic.py; never part of the product.
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <vector>
#include <algorithm>

namespace {
const uint8_t* g_base = nullptr; uint64_t g_bytes = 0;
std::vector<uint32_t> g_off;                 // field offsets, ascending
std::vector<uint64_t> g_bitmap;              // one bit per 32-byte sector of the slab, this window
std::vector<uint32_t> g_touched;             // sector indices set in this window
std::vector<uint64_t> g_loads, g_stores, g_sectors, g_rsect;     // per field, accumulated over windows (rsect: sectors first touched by a load)
std::vector<uint64_t> g_wfirst;              // bitmap: sector first touched by a store in this window
uint64_t g_windows = 0;
inline void touch(const void* p, unsigned n, bool store) {
  const uint8_t* a = (const uint8_t*)p;
  if (a < g_base || a >= g_base + g_bytes) return;
  const uint64_t o = (uint64_t)(a - g_base);
  const int f = (int)(std::upper_bound(g_off.begin(), g_off.end(), (uint32_t)o) - g_off.begin()) - 1;
  if (f < 0) return;
  (store ? g_stores : g_loads)[f]++;
  for (uint64_t s = o >> 5; s <= (o + n - 1) >> 5; ++s) {
    uint64_t& w = g_bitmap[s >> 6]; const uint64_t b = 1ull << (s & 63);
    if (!(w & b)) { w |= b; g_touched.push_back((uint32_t)s); if (store) g_wfirst[s >> 6] |= b; }
  }
}
}
extern "C" {
void __tsan_init() {}
void __tsan_func_entry(void*) {}
void __tsan_func_exit() {}
void __tsan_read1(void* p) { touch(p, 1, false); }
void __tsan_read2(void* p) { touch(p, 2, false); }
void __tsan_read4(void* p) { touch(p, 4, false); }
void __tsan_read8(void* p) { touch(p, 8, false); }
void __tsan_read16(void* p) { touch(p, 16, false); }
void __tsan_write1(void* p) { touch(p, 1, true); }
void __tsan_write2(void* p) { touch(p, 2, true); }
void __tsan_write4(void* p) { touch(p, 4, true); }
void __tsan_write8(void* p) { touch(p, 8, true); }
void __tsan_write16(void* p) { touch(p, 16, true); }
void __tsan_read_range(void* p, unsigned long n) { if (n) touch(p, (unsigned)n, false); }
void __tsan_write_range(void* p, unsigned long n) { if (n) touch(p, (unsigned)n, true); }

// called by emu.cpp (-DDCU_EMU_TRACE) around every window
void trace_begin_window(const uint8_t* base, uint64_t bytes, const uint32_t* off, int nfields) {
  if (g_base != base || g_bytes != bytes || (int)g_off.size() != nfields) {
    g_base = nullptr;                           // (no attribution while the tables are rebuilt)
    g_off.assign(off, off + nfields);
    g_bitmap.assign((bytes / 32 + 64) / 64 + 1, 0); g_wfirst.assign(g_bitmap.size(), 0);
    g_loads.assign(nfields, 0); g_stores.assign(nfields, 0); g_sectors.assign(nfields, 0); g_rsect.assign(nfields, 0);
    g_windows = 0; g_touched.clear();
    g_bytes = bytes; g_base = base;
  }
}
void trace_end_window() {
  const uint8_t* keep = g_base; g_base = nullptr;
  for (uint32_t s : g_touched) {
    const int f = (int)(std::upper_bound(g_off.begin(), g_off.end(), s * 32u) - g_off.begin()) - 1;
    if (f >= 0) { g_sectors[f]++; if (!(g_wfirst[s >> 6] & (1ull << (s & 63)))) g_rsect[f]++; }
    g_bitmap[s >> 6] = 0; g_wfirst[s >> 6] = 0;
  }
  g_touched.clear(); ++g_windows;
  g_base = keep;
}
// out[4 * f + {0,1,2,3}] = loads, stores, distinct sectors, sectors whose first access was a load (live across windows or uninitialised reads)
uint64_t trace_report(uint64_t* out, int cap) {
  for (int f = 0; f < (int)g_off.size() && f < cap; ++f) { out[4 * f] = g_loads[f]; out[4 * f + 1] = g_stores[f]; out[4 * f + 2] = g_sectors[f]; out[4 * f + 3] = g_rsect[f]; }
  return g_windows;
}
}
