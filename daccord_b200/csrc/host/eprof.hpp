// eprof.hpp -- error-profile estimation on the host: what daccord does once per .las before its main loop
// when no <las>.eprof exists (reference src/daccord.cpp:1652-1880 driver, :271-631 handleIndelEstimate<8>).
//
// Over the first <= 1024 A-reads: windows of 40 bases advancing by 5; the pile of a window is the A window plus the B
// slices of every selected overlap spanning it; piles of depth >= 3 in which no sequence repeats a 7-mer are "usable";
// their consensus is the single unitig of the k = 8 graph (k-mers seen >= 2 times) that leads from the most frequent
// first k-mer to the most frequent last k-mer; every pile sequence is aligned to it and the step counts (match, mismatch,
// insertion, deletion) summed over all windows are the profile the OffsetLikely / KmerLimit tables are built from.
//
// This runs once per input on a bounded sample, so it stays on the host (the product's hot path is the main loop).
// Own formulation: dense 4^8 count tables and explicit unitig vectors instead of the reference's sorted node arrays.
#pragma once
#include <cstdint>
#include <cmath>
#include <map>
#include <vector>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <stdexcept>
#include <sys/stat.h>
#include "las.hpp"
#include "synth.hpp"
#include "align.hpp"
#include "pile.hpp"

namespace dhost {

struct ProfileCounts {
  uint64_t cnt[4] = {0, 0, 0, 0};            // matches, mismatches, insertions, deletions
  uint64_t usable = 0, unusable = 0, reads = 0;
  double eavg = 0.0, edif = 0.0;             // mean / deviation of the per-read window error rate (reported only)
};

// <las>.eprof on disk, as the reference writes it (src/daccord.cpp:1855-1864): AlignmentStatistics::serialise = matches, mismatches,
// insertions, deletions as big-endian 64-bit numbers (libmaus2 NumberSerialisation::serialiseNumber), then eavg and edif by
// serialiseDouble (the 8 bytes of the double as they lie in memory; libmaus2 is not available here, so this follows its public
// sources).  The reference reads back only the four counts (:1864 GAS.deserialise).  The text form of this repository's first
// round ("m s i d" newline "eavg edif") is still accepted when reading.
inline void write_eprof(const std::string& fn, const uint64_t cnt[4], double eavg, double edif) {
  const std::string tmp = fn + ".tmp";                       // tmp + rename like the reference (:1854-1860)
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write error profile " + tmp);
  uint8_t b[48];
  for (int i = 0; i < 4; ++i) for (int q = 0; q < 8; ++q) b[8 * i + q] = (uint8_t)(cnt[i] >> (8 * (7 - q)));
  memcpy(b + 32, &eavg, 8); memcpy(b + 40, &edif, 8);
  const bool ok = fwrite(b, 1, sizeof(b), f) == sizeof(b);
  if (fclose(f) != 0 || !ok) throw std::runtime_error("cannot write error profile " + tmp);
  if (rename(tmp.c_str(), fn.c_str())) throw std::runtime_error("cannot rename " + tmp + " to " + fn);
}
inline bool read_eprof(const std::string& fn, uint64_t cnt[4]) {
  FILE* f = fopen(fn.c_str(), "rb");
  if (!f) return false;
  uint8_t b[256]; const size_t n = fread(b, 1, sizeof(b) - 1, f); fclose(f);
  if (n >= 32 && !(b[0] >= '0' && b[0] <= '9')) {            // binary (a count below 2^56 starts with a zero byte)
    for (int i = 0; i < 4; ++i) { cnt[i] = 0; for (int q = 0; q < 8; ++q) cnt[i] = (cnt[i] << 8) | b[8 * i + q]; }
    return true;
  }
  b[n] = 0;
  unsigned long long v[4];
  if (sscanf((const char*)b, "%llu %llu %llu %llu", &v[0], &v[1], &v[2], &v[3]) != 4) return false;
  for (int i = 0; i < 4; ++i) cnt[i] = v[i];
  return true;
}
// the reference recomputes the profile when the file is missing, or older than the .las unless --keepeprof (src/daccord.cpp:1652-1657)
inline bool eprof_is_stale(const std::string& eproffn, const std::string& lasfn, bool keepeprof) {
  struct stat se, sl;
  if (stat(eproffn.c_str(), &se) != 0) return true;
  if (keepeprof || stat(lasfn.c_str(), &sl) != 0) return false;
  return se.st_mtim.tv_sec < sl.st_mtim.tv_sec || (se.st_mtim.tv_sec == sl.st_mtim.tv_sec && se.st_mtim.tv_nsec < sl.st_mtim.tv_nsec);
}

struct Seq { const uint8_t* p; uint32_t n; };     // 2-bit codes, one per byte

// the k = 8 graph of one pile and its "trivial" consensus (reference DebruijnGraph::traverseTrivial, src/DebruijnGraph.hpp:3794-3826)
class TrivialConsensus {
 public:
  enum { K = 8, NK = 1 << (2 * K), MASK = NK - 1 };
  TrivialConsensus() : freq(NK, 0), mark(NK, 0) {}

  // returns false if the pile has no such unitig; cons = base codes
  bool run(const std::vector<Seq>& pile, std::vector<uint8_t>& cons) {
    cons.clear();
    for (uint32_t v : touched) freq[v] = 0;
    touched.clear(); firsts.clear(); lasts.clear();
    for (const Seq& s : pile) {
      if (s.n < K) continue;
      uint32_t v = 0;
      for (uint32_t i = 0; i < s.n; ++i) {
        v = ((v << 2) | s.p[i]) & MASK;
        if (i + 1 < K) continue;
        if (!freq[v]++) touched.push_back(v);
        if (i + 1 == K) firsts.push_back(v);
        if (i + 1 == s.n) lasts.push_back(v);
      }
    }
    // most frequent first k-mer among the graph's k-mers, most frequent last k-mer among all; ties: the smallest k-mer
    // (src/DebruijnGraph.hpp:1256-1278, :1329-1357); 0 if none
    const uint32_t F = most_frequent(firsts, true), L = most_frequent(lasts, false);
    if (!node(F) || !node(L)) return false;
    // unitigs leaving F (src/DebruijnGraph.hpp:2844-2986 without the predecessor check), cut where F or L occurs inside
    // (:2772-2841), one kept per first edge: the longest, then the one with the smaller end (:3087-3114)
    build_unitigs();
    split_at(F); split_at(L);
    const std::vector<uint32_t>* best = nullptr;
    std::map<uint32_t, const std::vector<uint32_t>*> by_ext;      // of the unitigs starting at F: first edge -> representative
    for (const auto& u : U) {
      if (u[0] != F) continue;
      const std::vector<uint32_t>*& rep = by_ext[u[1]];
      if (!rep || u.size() > rep->size() || (u.size() == rep->size() && u.back() < rep->back())) rep = &u;
    }
    for (auto& kv : by_ext) if (kv.second->back() == L) { best = kv.second; break; }
    if (!best) return false;
    for (int i = K - 1; i >= 0; --i) cons.push_back((uint8_t)((F >> (2 * i)) & 3));
    for (size_t j = 1; j < best->size(); ++j) cons.push_back((uint8_t)((*best)[j] & 3));
    return true;
  }

 private:
  std::vector<uint16_t> freq; std::vector<uint8_t> mark;
  std::vector<uint32_t> touched, firsts, lasts;
  std::vector<std::vector<uint32_t>> U;

  bool node(uint32_t v) const { return freq[v] >= 2; }                        // filterFreq(2), src/daccord.cpp:561
  uint32_t most_frequent(std::vector<uint32_t>& v, bool nodes_only) const {
    std::sort(v.begin(), v.end());
    uint32_t bestv = 0; size_t bestc = 0;
    for (size_t l = 0; l < v.size();) {
      size_t h = l; while (h < v.size() && v[h] == v[l]) ++h;
      if ((!nodes_only || node(v[l])) && h - l > bestc) { bestc = h - l; bestv = v[l]; }
      l = h;
    }
    return bestv;
  }
  // active out-edges of v: the most frequent successor and every other one with at least half its count, in order of
  // (count, symbol) descending (src/DebruijnGraph.hpp:1770-1814 with the k-mer limit switched off, :2434-2460)
  int active_succ(uint32_t v, uint32_t out[4]) const {
    uint32_t key[4]; int n = 0;
    for (uint32_t s = 0; s < 4; ++s) { uint32_t x = ((v << 2) & MASK) | s; if (node(x)) key[n++] = ((uint32_t)freq[x] << 8) | s; }
    std::sort(key, key + n, std::greater<uint32_t>());
    int na = 0;
    for (int i = 0; i < n; ++i) { if (i && (key[i] >> 8) < (key[0] >> 8) / 2) break; out[na++] = ((v << 2) & MASK) | (key[i] & 0xFF); }
    return na;
  }
  int active_pred_count(uint32_t v) const {                                   // src/DebruijnGraph.hpp:2552-2597
    int c = 0; uint32_t succ[4];
    for (uint32_t s = 0; s < 4; ++s) {
      uint32_t u = (v >> 2) | (s << (2 * (K - 1)));
      if (!node(u)) continue;
      int n = active_succ(u, succ);
      for (int i = 0; i < n; ++i) if (succ[i] == v) { ++c; break; }
    }
    return c;
  }
  void build_unitigs() {
    U.clear();
    std::vector<uint32_t> nodes;
    for (uint32_t v : touched) if (node(v)) nodes.push_back(v);
    uint32_t succ[4], nx[4];
    for (uint32_t v : nodes) {
      const int ns = active_succ(v, succ);
      if (!ns || (active_pred_count(v) == 1 && ns == 1)) continue;            // unitigs start at branch points and sources
      for (int e = 0; e < ns; ++e) {
        std::vector<uint32_t> u; u.push_back(v); u.push_back(succ[e]);
        mark[v] = 1; mark[succ[e]] = 1;
        bool loop = (v == succ[e]); uint32_t cur = succ[e];
        while (!loop && active_succ(cur, nx) == 1) {
          cur = nx[0]; u.push_back(cur);
          if (mark[cur]) loop = true; else mark[cur] = 1;
        }
        for (uint32_t x : u) mark[x] = 0;
        if (loop && cur != v) {          // ran into its own path: end at the first visit of that k-mer
          size_t j = 0; while (u[j] != cur) ++j;
          u.resize(j + 1);
        }
        U.push_back(std::move(u));
      }
    }
  }
  void split_at(uint32_t v) {
    const size_t n0 = U.size();
    std::vector<std::vector<uint32_t>> add; std::vector<uint8_t> drop(n0, 0);
    for (size_t z = 0; z < n0; ++z) {
      const auto& u = U[z];
      for (size_t i = 1; i + 1 < u.size(); ++i) if (u[i] == v) {
        add.emplace_back(u.begin(), u.begin() + i + 1); add.emplace_back(u.begin() + i, u.end()); drop[z] = 1; break;
      }
    }
    size_t o = 0;
    for (size_t z = 0; z < n0; ++z) if (!drop[z]) { if (o != z) U[o] = std::move(U[z]); ++o; }
    U.resize(o);
    for (auto& u : add) U.push_back(std::move(u));
  }
};

// does the sequence contain some k-mer twice (libmaus2::fastx::KmerRepeatDetector(k).detect, call site src/daccord.cpp:541)
inline bool repeats_kmer(const Seq& s, unsigned k, std::vector<uint32_t>& tmp) {
  if (s.n < k) return false;
  tmp.clear(); uint32_t v = 0; const uint32_t mask = (1u << (2 * k)) - 1;
  for (uint32_t i = 0; i < s.n; ++i) { v = ((v << 2) | s.p[i]) & mask; if (i + 1 >= k) tmp.push_back(v); }
  std::sort(tmp.begin(), tmp.end());
  for (size_t i = 1; i < tmp.size(); ++i) if (tmp[i] == tmp[i - 1]) return true;
  return false;
}

// keeps the D lowest-scoring overlaps of an A-read (the estimator's heap is ordered the other way round from the main
// loop's, src/daccord.cpp:1406-1412, :1707-1735), then orders by abpos, ties in file order
inline void select_overlaps_profile(const LasData& L, uint64_t aread, uint64_t maxinput, std::vector<uint32_t>& sel) {
  sel.clear();
  const uint64_t b = L.aidx[aread], e = L.aidx[aread + 1];
  if (e - b <= maxinput) { for (uint64_t i = b; i < e; ++i) sel.push_back((uint32_t)i); }
  else {
    // max-heap on score over slots, same array-heap procedure as ScoreHeap with the comparison reversed
    std::vector<std::pair<uint64_t, uint32_t>> H;
    auto before = [](const std::pair<uint64_t, uint32_t>& x, const std::pair<uint64_t, uint32_t>& y) { return x.first > y.first; };
    auto push = [&](std::pair<uint64_t, uint32_t> v) {
      H.push_back(v); size_t i = H.size() - 1;
      while (i > 0) { size_t p = (i - 1) >> 1; if (before(H[i], H[p])) { std::swap(H[i], H[p]); i = p; } else break; }
    };
    auto pop = [&]() {
      H[0] = H.back(); H.pop_back(); size_t p = 0, f = H.size();
      for (;;) { size_t l = 2 * p + 1, r = l + 1; if (l >= f) break; size_t m = (r < f && before(H[r], H[l])) ? r : l; if (before(H[m], H[p])) { std::swap(H[m], H[p]); p = m; } else break; }
    };
    for (uint64_t i = b; i < e; ++i) {
      const Overlap& o = L.ovl[i];
      const uint64_t score = (uint64_t)std::ldexp((double)o.diffs / (double)(o.aepos - o.abpos), 30);
      if (H.size() == maxinput) {
        if (score > H[0].first) continue;
        uint32_t slot = H[0].second; pop(); sel[slot] = (uint32_t)i; push({score, slot});
      } else { push({score, (uint32_t)sel.size()}); sel.push_back((uint32_t)i); }
    }
    std::sort(sel.begin(), sel.end());
  }
  std::stable_sort(sel.begin(), sel.end(), [&](uint32_t x, uint32_t y) { return L.ovl[x].abpos < L.ovl[y].abpos; });
}

class ProfileEstimator {
 public:
  enum { EW = 40, EA = 5 };                   // src/daccord.cpp:1279-1280
  ProfileEstimator(const PackedDB& rdb, const LasData& rL, uint64_t rmaxalign, uint64_t rmaxinput)
      : db(rdb), L(rL), maxalign(rmaxalign), maxinput(rmaxinput), RP(rdb, rL, PileParams()) {}

  // one A-read: adds to C, returns the mean error rate of its consensus windows (0 if none)
  double read(uint64_t aread, ProfileCounts& C) {
    select_overlaps_profile(L, aread, maxinput, sel);
    const size_t nintv = sel.size();
    if (!nintv) return 0.0;
    decode_read(db, (uint32_t)aread, false, RP.abuf);
    if (bm.size() < nintv) { bm.resize(nintv); bseq.resize(nintv); }
    uint64_t maxaepos = 0; double maxerate = 0.0, minerate = 1.0;
    for (size_t z = 0; z < nintv; ++z) {
      const Overlap& o = L.ovl[sel[z]];
      maxaepos = std::max<uint64_t>(maxaepos, (uint64_t)o.aepos);
      const double er = (double)o.diffs / (double)(o.aepos - o.abpos);
      maxerate = std::max(maxerate, er); minerate = std::min(minerate, er);
      RP.compute_bmap_exact(o, bm[z]);        // all traces up front (src/daccord.cpp:385-398)
      bseq[z] = RP.bbuf;                      // B read in the overlap's orientation
    }
    const double ediv = (maxerate > minerate) ? (maxerate - minerate) : 1.0;
    std::vector<uint64_t> key(nintv);
    for (size_t z = 0; z < nintv; ++z) {
      const Overlap& o = L.ovl[sel[z]];
      const double er = (double)o.diffs / (double)(o.aepos - o.abpos);
      key[z] = ((uint64_t)(((er - minerate) / ediv) * 4294967295.0) << 32) | z;
    }
    const uint64_t ylimit = (maxaepos + EA >= EW) ? (maxaepos + EA - EW) / EA : 0;
    std::map<uint64_t, size_t> active;        // pile order: (scaled error rate, index)
    size_t z = 0; double esum = 0.0; uint64_t ecnt = 0;
    for (uint64_t y = 0; y < ylimit; ++y) {
      const int64_t astart = (int64_t)(y * EA), aend = astart + EW;
      for (; z < nintv && L.ovl[sel[z]].abpos <= astart; ++z)
        if (L.ovl[sel[z]].aepos > aend) active[key[z]] = z;                   // joins at its first window, if it outlasts it (:440-485)
      for (auto it = active.begin(); it != active.end();) { if (L.ovl[sel[it->second]].aepos <= aend) it = active.erase(it); else ++it; }
      if (active.empty()) continue;
      pile.clear();
      pile.push_back(Seq{RP.abuf.data() + astart, EW});
      for (auto& kv : active) {
        if (pile.size() >= maxalign) break;
        const Overlap& o = L.ovl[sel[kv.second]];
        const std::vector<uint32_t>& m = bm[kv.second];
        const size_t n0 = (size_t)(astart - o.abpos);
        pile.push_back(Seq{bseq[kv.second].data() + o.bbpos + m[n0], m[n0 + EW] - m[n0]});
      }
      if (pile.size() < 3) continue;
      bool rep = false;
      for (const Seq& s : pile) if (repeats_kmer(s, TrivialConsensus::K - 1, tmp)) { rep = true; break; }
      if (rep) { C.unusable++; continue; }
      C.usable++;
      if (!TC.run(pile, cons)) continue;
      uint64_t g[4] = {0, 0, 0, 0};
      for (const Seq& s : pile) RP.TA.align_count(cons.data(), (int)cons.size(), s.p, (int)s.n, g);
      for (int i = 0; i < 4; ++i) C.cnt[i] += g[i];
      esum += (double)(g[1] + g[2] + g[3]) / (double)(g[0] + g[1] + g[2] + g[3]);
      ++ecnt;
    }
    return ecnt ? esum / (double)ecnt : 0.0;
  }

 private:
  const PackedDB& db; const LasData& L; uint64_t maxalign, maxinput;
  ReadPiler RP;                                // for its tile aligner and trace reconstruction
  TrivialConsensus TC;
  std::vector<uint32_t> sel, tmp;
  std::vector<std::vector<uint32_t>> bm; std::vector<std::vector<uint8_t>> bseq;
  std::vector<Seq> pile; std::vector<uint8_t> cons;
};

// A-reads [lo, min(hi, lo + 1024)); per-read results are combined in read order whatever the thread count
inline ProfileCounts estimate_profile(const PackedDB& db, const LasData& L, int64_t lo, int64_t hi, uint64_t maxalign, uint64_t maxinput, int nthreads) {
  const int64_t top = std::min(hi, lo + 1024), nr = std::max<int64_t>(top - lo, 0);
  std::vector<ProfileCounts> per(nr); std::vector<double> rate(nr, 0.0);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    ProfileEstimator PE(db, L, maxalign, maxinput);
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < nr; ++i) {
      if (L.aidx[lo + i] == L.aidx[lo + i + 1]) continue;
      per[i].reads = 1;
      rate[i] = PE.read((uint64_t)(lo + i), per[i]);
    }
  }
  ProfileCounts R; double s = 0.0; uint64_t n = 0;
  for (int64_t i = 0; i < nr; ++i) {
    for (int j = 0; j < 4; ++j) R.cnt[j] += per[i].cnt[j];
    R.usable += per[i].usable; R.unusable += per[i].unusable; R.reads += per[i].reads;
    if (rate[i] != 0.0) { s += rate[i]; ++n; }
  }
  if (n) {
    R.eavg = s / (double)n; double d = 0.0;
    for (int64_t i = 0; i < nr; ++i) if (rate[i] != 0.0) d += (R.eavg - rate[i]) * (R.eavg - rate[i]);
    R.edif = std::sqrt(d / (double)n);
  }
  return R;
}

}  // namespace dhost
