// TEST INFRASTRUCTURE (oracle) -- parity unpinned, see oracle/README.md.
// extern "C" entry points so tests / smoke() / bench.py's cpu_baseline leg can drive the oracle
// on exactly the batches the product consumes (same dcu_window / dcu_slice / packed-DB inputs).
#include "window.hpp"
#include "../include/daccord_b200.h"
#include <cstring>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
using namespace oracle;

static Params to_params(const dcu_params* p) {
  Params P; P.w = p->w; P.k_lo = p->k_lo; P.k_hi = p->k_hi; P.minwindowcov = p->min_cov; P.minfilterfreq = p->min_ff; P.maxfilterfreq = p->max_ff;
  P.eminrate = p->max_err; P.p_i = p->p_i; P.p_d = p->p_d; P.est_cor = p->est_cor; return P;
}

extern "C" {

// decode a slice to ASCII (Dazzler .bps packing, first base in the top two bits; rc per dcu_slice.flags)
static void decode_slice(const uint8_t* packed, const dcu_slice& s, std::string& out) {
  out.resize(s.len);
  for (uint32_t i = 0; i < s.len; ++i) {
    uint32_t g = (s.flags & 1) ? s.gpos + (s.len - 1 - i) : s.gpos + i;
    unsigned c = (packed[g >> 2] >> (6 - 2 * (g & 3))) & 3;
    if (s.flags & 1) c = 3 - c;
    out[i] = "ACGT"[c];
  }
}

// returns elapsed seconds of the window loop (tables excluded), or a negative value on error
double oracle_run_batch(const dcu_params* prm, const uint8_t* packed, const dcu_window* win, uint64_t nwin, const dcu_slice* sl,
                        dcu_result* res, uint8_t* cons, uint8_t* ops, int nthreads) {
  try {
    Tables T(to_params(prm));
    if (nthreads < 1) nthreads = 1;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(nthreads)
    {
      WindowContext C(T);
      std::vector<std::string> S; std::vector<SeqRef> MA;
#pragma omp for schedule(dynamic, 64)
      for (int64_t i = 0; i < (int64_t)nwin; ++i) {
        const dcu_window& W = win[i];
        S.resize(W.slice_cnt); MA.resize(W.slice_cnt);
        for (uint32_t j = 0; j < W.slice_cnt; ++j) { decode_slice(packed, sl[W.slice_begin + j], S[j]); MA[j] = {(const uint8_t*)S[j].data(), S[j].size()}; }
        WindowResult R = C.run(MA.data(), MA.size());
        dcu_result& r = res[i];
        memset(&r, 0, sizeof(r));
        r.status = !R.attempted ? DCU_WIN_SKIPPED : (R.ok ? DCU_WIN_OK : DCU_WIN_FAILED);
        r.elength = (int32_t)R.elength; r.ff = -1;
        if (R.ok) {
          r.k = (uint8_t)R.k; r.ff = (int8_t)R.filterfreq; r.clen = (uint8_t)R.cons.size(); r.err = (uint32_t)R.minrate;
          r.nops = (uint16_t)R.trace.size(); r.ncand = (uint16_t)R.ncand;
          if (cons) memcpy(cons + i * DCU_CONS_STRIDE, R.cons.data(), std::min<size_t>(R.cons.size(), DCU_CONS_STRIDE));
          if (ops) memcpy(ops + i * DCU_OPS_STRIDE, R.trace.data(), std::min<size_t>(R.trace.size(), DCU_OPS_STRIDE));
        }
      }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  } catch (...) { return -1.0; }
}

// dump the oracle's tables in the product's dense layout: which = 0 DPnorm, 1 DPnormSquare, 2 VS (as double),
// 3 support lo/hi interleaved, 4 KmerLimit rows (KLIMN columns). returns count (needed size if cap too small)
int64_t oracle_get_tables(const dcu_params* prm, int which, double* out, int64_t cap, int klimn) {
  Tables T(to_params(prm));
  const OffsetLikely& OL = T.OL;
  int64_t NP = (int64_t)OL.DP.size(), MS = (int64_t)OL.Vsupport.size();
  std::vector<double> v;
  if (which <= 2) {
    v.assign(NP * MS, 0.0);
    for (int64_t l = 0; l < NP; ++l) {
      const DotVec& d = which == 0 ? OL.DPnorm[l] : OL.DPnormSquare[l];
      for (size_t o = 0; o < d.V.size(); ++o) v[l * MS + d.firstsign + o] = which == 2 ? (double)d.VS[o] : d.V[o];
    }
  } else if (which == 3) {
    for (int64_t i = 0; i < MS; ++i) { v.push_back((double)OL.Vsupport[i].first); v.push_back((double)OL.Vsupport[i].second); }
  } else if (which == 4) {
    for (auto& kv : T.MKL) { KmerLimit KL = kv.second; for (int n = 0; n < klimn; ++n) v.push_back(KL.getLimit(n)); }
  } else if (which == 5) { v.push_back((double)NP); v.push_back((double)MS); }
  if ((int64_t)v.size() <= cap) memcpy(out, v.data(), v.size() * sizeof(double));
  return (int64_t)v.size();
}

}  // extern "C"

// ---- read-level driver: .las + Dazzler DB in, FastA text out (the daccord CLI surface on the CPU)
#include "pipeline.hpp"
extern "C" {
// processes A-reads [first, last] (inclusive, reference -I semantics); returns malloc'ed FastA text
char* oracle_daccord_files(const dcu_params* prm, uint32_t advance, uint64_t maxalign, uint64_t maxinput, int producefull, uint64_t minlen,
                           const char* lasfn, const char* dbfn, int64_t first, int64_t last, int nthreads, uint64_t* outlen, uint64_t* stats) {
  try {
    Params P = to_params(prm); P.a = advance; P.maxalign = maxalign; P.maxinput = maxinput; P.producefull = producefull != 0; P.minlen = minlen;
    Tables T(P);
    ReadDB DB; loadDB(dbfn, DB);
    LasFile L; loadLas(lasfn, L, DB.reads.size());
    if (first < 0) first = 0;
    if (last < 0 || last >= (int64_t)DB.reads.size()) last = (int64_t)DB.reads.size() - 1;
    const int64_t nr = last >= first ? last - first + 1 : 0;
    std::vector<std::string> parts(nr);
    std::vector<ReadStats> st(nr);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
      ReadHandler H(T, L, DB);
#pragma omp for schedule(dynamic, 1)
      for (int64_t i = 0; i < nr; ++i) { uint64_t c = 0; st[i] = H.handle((uint64_t)(first + i), c, parts[i]); }
    }
    std::string out; uint64_t counter = 0;
    uint64_t s0 = 0, s1 = 0, s2 = 0;
    for (int64_t i = 0; i < nr; ++i) {
      s0 += st[i].windows; s1 += st[i].attempted; s2 += st[i].ok;
      std::istringstream is(parts[i]); std::string line;
      while (std::getline(is, line)) {
        if (!line.empty() && line[0] == '>') { size_t a = line.find('/'), b = line.find('/', a + 1); line = line.substr(0, a + 1) + std::to_string(counter++) + line.substr(b); }
        out += line; out.push_back('\n');
      }
    }
    if (stats) { stats[0] = s0; stats[1] = s1; stats[2] = s2; }
    char* buf = (char*)malloc(out.size() + 1); memcpy(buf, out.data(), out.size()); buf[out.size()] = 0;
    *outlen = out.size();
    return buf;
  } catch (std::exception& e) { fprintf(stderr, "[oracle] %s\n", e.what()); return nullptr; }
}
// error profile of A-reads [first, top) (at most 1024 are used): out = matches, mismatches, insertions, deletions, usable, unusable, reads;
// dout = eavg, edif
int oracle_estimate_profile(const char* lasfn, const char* dbfn, int64_t first, int64_t top, uint64_t maxalign, uint64_t maxinput, uint64_t* out, double* dout) {
  try {
    ReadDB DB; loadDB(dbfn, DB);
    LasFile L; loadLas(lasfn, L, DB.reads.size());
    if (first < 0) first = 0;
    if (top < 0 || top > (int64_t)DB.reads.size()) top = (int64_t)DB.reads.size();
    ProfileResult R = estimateProfile(L, DB, first, top, maxalign, maxinput);
    out[0] = R.stats.matches; out[1] = R.stats.mismatches; out[2] = R.stats.insertions; out[3] = R.stats.deletions; out[4] = R.usable; out[5] = R.unusable; out[6] = R.nreads;
    if (dout) { dout[0] = R.eavg; dout[1] = R.edif; }
    return 0;
  } catch (std::exception& e) { fprintf(stderr, "[oracle] %s\n", e.what()); return 1; }
}
void oracle_free(void* p) { free(p); }
// both edit distances of one pair (tests: the bit-parallel scoring distance equals the DP definition)
void oracle_edit_distances(const uint8_t* a, uint64_t la, const uint8_t* b, uint64_t lb, uint64_t* out) {
  out[0] = oracle::editDistance(a, la, b, lb); out[1] = oracle::editDistanceDP(a, la, b, lb);
}
}
