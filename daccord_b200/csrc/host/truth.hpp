// truth.hpp -- accuracy of corrected reads against the simulated truth (tests and bench.py only; an anchor outside the oracle:
// the reference quotes the error rate of its output against the true sequence, README.md:442).  The simulator keeps the genome and
// every read's edit script (synth.hpp); a corrected segment "A=[first,last]" of read r is compared with the genome interval
// those A positions were sampled from.
#pragma once
#include "synth.hpp"
#include "simlas.hpp"
#include <string>
#include <vector>
#include <cstdlib>

namespace dhost {

struct TruthRead { uint64_t gstart = 0, glen = 0; bool rc = false; uint32_t len = 0; std::vector<int16_t> drift; std::vector<uint16_t> cerr; };
struct Truth {
  std::vector<uint8_t> genome; std::vector<TruthRead> reads;
  bool have() const { return !reads.empty(); }
};
inline void keep_truth(const std::vector<uint8_t>& G, std::vector<SimRead>& reads, Truth& T) {
  T.genome = G; T.reads.resize(reads.size());
  for (size_t i = 0; i < reads.size(); ++i) {
    TruthRead& t = T.reads[i]; SimRead& r = reads[i];
    t.gstart = r.gstart; t.glen = r.glen; t.rc = r.rc; t.len = (uint32_t)r.seq.size(); t.drift.swap(r.drift); t.cerr.swap(r.cerr);
  }
}
// unit-cost global edit distance inside a diagonal band (half width W around the corner-to-corner diagonal): exact when the optimal
// path stays inside, otherwise an upper bound
inline uint64_t banded_distance(const std::string& a, const std::string& b, int64_t W) {
  const int64_t n = (int64_t)a.size(), m = (int64_t)b.size();
  if (!n) return (uint64_t)m;
  if (!m) return (uint64_t)n;
  const int64_t d = m - n; W += std::llabs(d);
  const int32_t INF = 1 << 29;
  std::vector<int32_t> prev(2 * W + 3, INF), cur(2 * W + 3, INF);      // cell (i, j) at index j - i + W + 1
  for (int64_t j = 0; j <= std::min(m, W); ++j) prev[j + W + 1] = (int32_t)j;
  for (int64_t i = 1; i <= n; ++i) {
    const int64_t jlo = std::max<int64_t>(0, i - W), jhi = std::min(m, i + W);
    std::fill(cur.begin(), cur.end(), INF);
    for (int64_t j = jlo; j <= jhi; ++j) {
      const int64_t x = j - i + W + 1;
      int32_t v = INF;
      if (j == 0) v = (int32_t)i;
      else {
        v = prev[x] + (a[i - 1] != b[j - 1]);                        // (i-1, j-1) sits at the same band index of the previous row
        if (cur[x - 1] + 1 < v) v = cur[x - 1] + 1;                   // (i, j-1)
      }
      if (prev[x + 1] + 1 < v) v = prev[x + 1] + 1;                   // (i-1, j)
      cur[x] = v;
    }
    prev.swap(cur);
  }
  return (uint64_t)prev[m - n + W + 1];
}
struct TruthStats { uint64_t segments = 0, bases = 0, truth_bases = 0, edits = 0, raw_events = 0, reads = 0; };
// every segment of a FastA text (headers ">id/counter/first_end A=[first,last]", 80-column lines) whose read id is < max_read
inline bool truth_eval(const Truth& T, const char* fasta, uint64_t len, uint64_t max_read, TruthStats& S, std::string& err) {
  struct Seg { uint64_t read, first, last; std::string s; };
  std::vector<Seg> segs;
  uint64_t p = 0;
  while (p < len) {
    uint64_t e = p; while (e < len && fasta[e] != '\n') ++e;
    if (fasta[p] == '>') {
      unsigned long long id = 0, c = 0, f = 0, g = 0, a0 = 0, a1 = 0;
      if (sscanf(std::string(fasta + p, e - p).c_str(), ">%llu/%llu/%llu_%llu A=[%llu,%llu]", &id, &c, &f, &g, &a0, &a1) != 6 || id == 0) { err = "unparsable FastA header"; return false; }
      segs.push_back(Seg{id - 1, a0, a1, std::string()});
    } else if (!segs.empty()) {
      for (uint64_t i = p; i < e; ++i) { char ch = fasta[i]; if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32); segs.back().s.push_back(ch); }
    }
    p = e + 1;
  }
  uint64_t segments = 0, bases = 0, tb = 0, edits = 0, raw = 0; int bad = 0;
  std::vector<uint8_t> seen(T.reads.size(), 0);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : segments, bases, tb, edits, raw) reduction(max : bad)
  for (int64_t i = 0; i < (int64_t)segs.size(); ++i) {
    const Seg& g = segs[i];
    if (g.read >= max_read) continue;
    if (g.read >= T.reads.size()) { bad = 1; continue; }
    const TruthRead& R = T.reads[g.read];
    if (g.last < g.first || g.last >= R.len) { bad = 1; continue; }
    // read positions -> forward copy indices -> genome offsets
    const uint64_t f0 = R.rc ? (uint64_t)R.len - 1 - g.last : g.first, f1 = R.rc ? (uint64_t)R.len - 1 - g.first : g.last;
    auto goff = [&](uint64_t fi) { uint64_t lo = 0, hi = R.glen; while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if ((uint64_t)((int64_t)mid + R.drift[mid]) >= fi) hi = mid; else lo = mid + 1; } return lo; };
    const uint64_t o0 = goff(f0), o1 = goff(f1 + 1);
    std::string t; t.reserve(o1 - o0);
    if (!R.rc) for (uint64_t o = o0; o < o1; ++o) t.push_back("ACGT"[T.genome[R.gstart + o]]);
    else for (uint64_t o = o1; o > o0; --o) t.push_back("ACGT"[3 - T.genome[R.gstart + o - 1]]);
    edits += banded_distance(g.s, t, 64);
    raw += (uint64_t)(R.cerr[o1] - R.cerr[o0]);
    segments += 1; bases += g.s.size(); tb += t.size();
    seen[g.read] = 1;
  }
  if (bad) { err = "segment outside its read"; return false; }
  S.segments = segments; S.bases = bases; S.truth_bases = tb; S.edits = edits; S.raw_events = raw;
  for (auto v : seen) S.reads += v;
  return true;
}

}  // namespace dhost
