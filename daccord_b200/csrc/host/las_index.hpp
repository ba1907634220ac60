// las_index.hpp -- ranged, parallel ingest of DALIGNER .las files.
// The reference never reads a whole .las: it builds an index of record offsets per A-read once (rebuilt when older than the
// .las, reference src/daccord.cpp:1075-1104), takes the A-read range of the file from it (:1106-1117) and opens a file region
// per A-read (:2129-2140, OverlapIndexer::openAlignmentFileRegion).  Same idea here, own format:
//   build_las_index : one sequential pass over the record headers of the memory-mapped file -> byte offset of the first record of
//                     every A-read; cached next to the .las as <las>.dcuidx (validated by file size and mtime)
//   read_las_range  : maps only the byte range of A-reads [lo, hi), finds the record boundaries, then decodes the records on
//                     all host threads -- what a -I / -J shard (one process per GPU) needs instead of the whole file
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "las.hpp"

namespace dhost {

struct LasIndex {
  int32_t tspace = 0; int64_t novl = 0;
  int64_t minaread = 0, maxaread = -1;          // A-read ids present in the file (maxaread < minaread: empty file)
  uint64_t file_size = 0; int64_t mtime_ns = 0;
  std::vector<uint64_t> off;                    // off[i] = byte offset of the first record with aread >= minaread + i; off.back() = file size
  std::vector<uint64_t> rec;                    // rec[i] = number of records before that offset
  // byte range / record range of A-reads [lo, hi)
  void range(int64_t lo, int64_t hi, uint64_t& b0, uint64_t& b1, uint64_t& r0, uint64_t& r1) const {
    const int64_t n = (int64_t)off.size() - 1;
    int64_t i0 = std::min<int64_t>(std::max<int64_t>(lo - minaread, 0), n), i1 = std::min<int64_t>(std::max<int64_t>(hi - minaread, i0), n);
    b0 = off[(size_t)i0]; b1 = off[(size_t)i1]; r0 = rec[(size_t)i0]; r1 = rec[(size_t)i1];
  }
};

struct MappedFile {
  const uint8_t* p = nullptr; size_t map_len = 0; uint64_t size = 0; int64_t mtime_ns = 0; size_t delta = 0; void* base = nullptr;
  // maps bytes [b0, b1) of the file (the whole file if b1 == 0); p points at byte b0
  void open(const std::string& fn, uint64_t b0 = 0, uint64_t b1 = 0) {
    int fd = ::open(fn.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + fn);
    struct stat st;
    if (fstat(fd, &st)) { ::close(fd); throw std::runtime_error("cannot stat " + fn); }
    size = (uint64_t)st.st_size; mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    if (b1 == 0 || b1 > size) b1 = size;
    if (b0 > b1) b0 = b1;
    const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE), a0 = b0 / page * page;
    delta = (size_t)(b0 - a0); map_len = (size_t)(b1 - a0);
    if (map_len) {
      base = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, (off_t)a0);
      if (base == MAP_FAILED) { base = nullptr; ::close(fd); throw std::runtime_error("cannot map " + fn); }
      madvise(base, map_len, MADV_SEQUENTIAL);
      p = (const uint8_t*)base + delta;
    }
    ::close(fd);
  }
  ~MappedFile() { if (base) munmap(base, map_len); }
};

inline size_t las_trace_bytes(int32_t tspace) { return tspace <= 125 ? 1 : 2; }

inline void build_las_index(const std::string& fn, LasIndex& I) {
  MappedFile M; M.open(fn);
  if (M.size < 12) throw std::runtime_error("short LAS header in " + fn);
  memcpy(&I.novl, M.p, 8); memcpy(&I.tspace, M.p + 8, 4);
  I.file_size = M.size; I.mtime_ns = M.mtime_ns; I.off.clear(); I.rec.clear(); I.minaread = 0; I.maxaread = -1;
  const size_t tb = las_trace_bytes(I.tspace);
  uint64_t pos = 12; int64_t prev = -1;
  for (int64_t r = 0; r < I.novl; ++r) {
    if (pos + 40 > M.size) throw std::runtime_error("truncated LAS record in " + fn);
    int32_t tlen, aread; memcpy(&tlen, M.p + pos, 4); memcpy(&aread, M.p + pos + 28, 4);
    if (tlen < 0 || aread < 0) throw std::runtime_error("bad LAS record in " + fn);
    if (prev < 0) { I.minaread = aread; prev = aread - 1; }
    if (aread < prev) throw std::runtime_error("LAS file is not sorted by A-read: " + fn);
    for (; prev < aread; ++prev) { I.off.push_back(pos); I.rec.push_back((uint64_t)r); }     // reads without overlaps share the next offset
    I.maxaread = aread;
    pos += 40 + (uint64_t)tlen * tb;
  }
  if (pos > M.size) throw std::runtime_error("truncated LAS trace in " + fn);
  I.off.push_back(pos); I.rec.push_back((uint64_t)I.novl);
}

// cache file: magic, tspace, novl, minaread, maxaread, file_size, mtime_ns, n, then off[n], rec[n]
inline bool load_las_index(const std::string& fn, LasIndex& I) {
  struct stat st;
  if (stat(fn.c_str(), &st)) return false;
  FILE* f = fopen((fn + ".dcuidx").c_str(), "rb");
  if (!f) return false;
  uint64_t magic = 0, n = 0; int64_t hdr[6]; bool ok = false;
  if (fread(&magic, 8, 1, f) == 1 && magic == 0x3158444955434441ull && fread(hdr, 8, 6, f) == 6 && fread(&n, 8, 1, f) == 1 && n >= 1 && n < (1ull << 40)) {
    I.tspace = (int32_t)hdr[0]; I.novl = hdr[1]; I.minaread = hdr[2]; I.maxaread = hdr[3]; I.file_size = (uint64_t)hdr[4]; I.mtime_ns = hdr[5];
    const int64_t mt = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    if (I.file_size == (uint64_t)st.st_size && I.mtime_ns == mt) {
      I.off.resize(n); I.rec.resize(n);
      ok = fread(I.off.data(), 8, n, f) == n && fread(I.rec.data(), 8, n, f) == n && I.off.back() <= I.file_size;
    }
  }
  fclose(f);
  return ok;
}
inline bool save_las_index(const std::string& fn, const LasIndex& I) {      // best effort, via tmp + rename
  const std::string tmp = fn + ".dcuidx.tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return false;
  const uint64_t magic = 0x3158444955434441ull, n = I.off.size();
  const int64_t hdr[6] = {I.tspace, I.novl, I.minaread, I.maxaread, (int64_t)I.file_size, I.mtime_ns};
  bool ok = fwrite(&magic, 8, 1, f) == 1 && fwrite(hdr, 8, 6, f) == 6 && fwrite(&n, 8, 1, f) == 1 && fwrite(I.off.data(), 8, n, f) == n && fwrite(I.rec.data(), 8, n, f) == n;
  ok = (fclose(f) == 0) && ok;
  if (ok) ok = rename(tmp.c_str(), (fn + ".dcuidx").c_str()) == 0;
  if (!ok) remove(tmp.c_str());
  return ok;
}
// the index of a .las: cached copy if it matches the file, else built (and cached when the directory is writable)
inline void get_las_index(const std::string& fn, LasIndex& I, bool* built = nullptr) {
  if (load_las_index(fn, I)) { if (built) *built = false; return; }
  build_las_index(fn, I);
  save_las_index(fn, I);
  if (built) *built = true;
}

// overlaps of A-reads [lo, hi) only; record boundaries sequentially (headers only), record bodies on nthreads threads
inline void read_las_range(const std::string& fn, const LasIndex& I, int64_t lo, int64_t hi, LasData& L, int nthreads = 1) {
  uint64_t b0, b1, r0, r1; I.range(lo, hi, b0, b1, r0, r1);
  L.tspace = I.tspace; L.ovl.clear(); L.trace.clear();
  const uint64_t n = r1 - r0;
  if (!n) return;
  MappedFile M; M.open(fn, b0, b1);
  if (M.size != I.file_size) throw std::runtime_error("LAS file changed since it was indexed: " + fn);
  const size_t tb = las_trace_bytes(I.tspace); const uint64_t len = b1 - b0;
  std::vector<uint64_t> rpos(n + 1), tpos(n + 1);
  uint64_t pos = 0, tv = 0;
  for (uint64_t r = 0; r < n; ++r) {
    if (pos + 40 > len) throw std::runtime_error("truncated LAS record in " + fn);
    int32_t tlen; memcpy(&tlen, M.p + pos, 4);
    if (tlen < 0) throw std::runtime_error("bad trace length in " + fn);
    rpos[r] = pos; tpos[r] = tv; pos += 40 + (uint64_t)tlen * tb; tv += (uint64_t)tlen;
  }
  if (pos != len) throw std::runtime_error("LAS index does not match the file: " + fn);
  rpos[n] = pos; tpos[n] = tv;
  L.ovl.resize(n); L.trace.resize(tv);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t r = 0; r < (int64_t)n; ++r) {
    int32_t rec[10]; memcpy(rec, M.p + rpos[r], 40);
    Overlap& o = L.ovl[(size_t)r];
    o.tlen = rec[0]; o.diffs = rec[1]; o.abpos = rec[2]; o.bbpos = rec[3]; o.aepos = rec[4]; o.bepos = rec[5];
    o.flags = (uint32_t)rec[6]; o.aread = rec[7]; o.bread = rec[8]; o.trace_off = tpos[r];
    const uint8_t* t = M.p + rpos[r] + 40; uint16_t* d = L.trace.data() + tpos[r];
    if (tb == 1) for (int32_t k = 0; k < o.tlen; ++k) d[k] = t[k];
    else memcpy(d, t, (size_t)o.tlen * 2);
  }
}

}  // namespace dhost
