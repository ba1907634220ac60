// TEST INFRASTRUCTURE (oracle) -- parity unpinned, see oracle/README.md.
// Read-level driver of daccord restated: overlap selection (reference src/daccord.cpp:2112-2288), the window loop
// of HandleContext::operator() (src/HandleContext.hpp:1699-2901) with trace reconstruction (convention C8), pile
// vote and FastA output (:2541-2724), plus minimal .las / Dazzler-DB readers (formats per SURVEY.md appendix B).
// Independent of daccord_b200/csrc/host: different containers, explicit step strings (advanceA /
// getStringLengthUsed exactly as the reference uses them) instead of the product's precomputed offset maps.
#pragma once
#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <sstream>
#include <fstream>
#include <stdexcept>
#include "window.hpp"

namespace oracle {

// ------------------------------------------------------------------ file formats
struct OvlRec { int32_t tlen, diffs, abpos, bbpos, aepos, bepos; uint32_t flags; int32_t aread, bread; std::vector<uint16_t> tr; };
struct LasFile { int32_t tspace = 0; std::vector<OvlRec> ovl; std::vector<uint64_t> first; };   // first[r] = index of first overlap of A-read r
struct ReadDB { std::vector<std::string> reads; };

inline void loadLas(const std::string& fn, LasFile& L, uint64_t nreads) {
  std::ifstream f(fn, std::ios::binary);
  if (!f) throw std::runtime_error("oracle: cannot open " + fn);
  int64_t novl = 0; f.read((char*)&novl, 8); f.read((char*)&L.tspace, 4);
  L.ovl.resize((size_t)novl);
  for (auto& o : L.ovl) {
    int32_t rec[10]; f.read((char*)rec, 40);
    if (!f) throw std::runtime_error("oracle: truncated las");
    o.tlen = rec[0]; o.diffs = rec[1]; o.abpos = rec[2]; o.bbpos = rec[3]; o.aepos = rec[4]; o.bepos = rec[5]; o.flags = (uint32_t)rec[6]; o.aread = rec[7]; o.bread = rec[8];
    o.tr.resize(o.tlen);
    if (L.tspace <= 125) { std::vector<uint8_t> b(o.tlen); f.read((char*)b.data(), o.tlen); for (int i = 0; i < o.tlen; ++i) o.tr[i] = b[i]; }
    else f.read((char*)o.tr.data(), 2 * o.tlen);
  }
  L.first.assign(nreads + 1, 0);
  for (auto& o : L.ovl) L.first[(size_t)o.aread + 1]++;
  for (uint64_t i = 0; i < nreads; ++i) L.first[i + 1] += L.first[i];
}
inline std::string hiddenName(const std::string& db, const char* ext) {
  std::string dir, base = db; size_t sl = db.find_last_of('/');
  if (sl != std::string::npos) { dir = db.substr(0, sl + 1); base = db.substr(sl + 1); }
  if (base.size() > 3 && base.substr(base.size() - 3) == ".db") base.resize(base.size() - 3);
  return dir + "." + base + ext;
}
inline void loadDB(const std::string& dbfn, ReadDB& D) {
  std::ifstream fi(hiddenName(dbfn, ".idx"), std::ios::binary), fb(hiddenName(dbfn, ".bps"), std::ios::binary);
  if (!fi || !fb) throw std::runtime_error("oracle: cannot open Dazzler DB " + dbfn);
  // DAZZ_DB header: ureads,treads,cutoff,allarr (4x int32) freq[4] (float) maxlen (int32) +pad totlen (int64) nreads,trimmed,part,ufirst,tfirst (5x int32) +pad, 5 pointers/ints
  unsigned char hdr[112]; fi.read((char*)hdr, sizeof(hdr));
  int32_t nreads; memcpy(&nreads, hdr + 48, 4);
  std::vector<char> bps((std::istreambuf_iterator<char>(fb)), std::istreambuf_iterator<char>());
  D.reads.resize(nreads);
  for (int32_t i = 0; i < nreads; ++i) {
    unsigned char rec[40]; fi.read((char*)rec, sizeof(rec));          // DAZZ_READ: origin,rlen,fpulse (int32) pad boff,coff (int64) flags (int32) pad
    int32_t rlen; int64_t boff; memcpy(&rlen, rec + 4, 4); memcpy(&boff, rec + 16, 8);
    std::string& s = D.reads[i]; s.resize(rlen);
    for (int32_t j = 0; j < rlen; ++j) s[j] = "ACGT"[((unsigned char)bps[boff + (j >> 2)] >> (6 - 2 * (j & 3))) & 3];
  }
}
inline std::string revcomp(const std::string& s) {
  std::string r(s.size(), 'A');
  for (size_t i = 0; i < s.size(); ++i) { char c = s[s.size() - 1 - i]; r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }
  return r;
}

// ------------------------------------------------------------------ one A-read
struct PileElement {            // src/HandleContext.hpp:219-248
  int64_t apos, apre; char sym;
  bool operator<(const PileElement& P) const { return apos != P.apos ? apos < P.apos : (apre != P.apre ? apre < P.apre : sym < P.sym); }
};
struct WindowsOf {              // src/HandleContext.hpp:382-447
  uint64_t l, a, w, n;
  static uint64_t computeN(uint64_t l, uint64_t a, uint64_t w) {
    uint64_t npre = (l + a >= w) ? ((l + a - w) / a) : 0;
    if (npre) return ((npre - 1) * a + w == l) ? npre : npre + 1;
    return l >= w ? 1 : 0;
  }
  WindowsOf(uint64_t rl, uint64_t ra, uint64_t rw) : l(rl), a(ra), w(rw), n(computeN(rl, ra, rw)) {}
  std::pair<uint64_t, uint64_t> at(uint64_t i) const { return (i * a + w <= l) ? std::make_pair(i * a, i * a + w) : std::make_pair(l - w, l); }
  uint64_t offset(uint64_t i) const { return (i + 1 < n) ? at(i + 1).first - at(i).first : 0; }
};

struct ActiveElement { const uint8_t* ua; const uint8_t* ub; const uint8_t* ta; const uint8_t* te; uint64_t uboff; };   // src/ActiveElement.hpp:26-49

struct ReadStats { uint64_t windows = 0, attempted = 0, ok = 0; };

struct ReadHandler {
  const Tables& T; WindowContext WC; const LasFile& L; const ReadDB& DB;
  std::map<int64_t, std::string> rccache;
  ReadHandler(const Tables& rT, const LasFile& rL, const ReadDB& rDB) : T(rT), WC(rT), L(rL), DB(rDB) {}

  // convention C8: trace = concatenation of per-tile global alignments (A tile vs the B block of the stated length)
  void computeTrace(const OvlRec& o, const uint8_t* a, const uint8_t* b, std::vector<uint8_t>& trace) {
    trace.clear();
    int64_t x = o.abpos, bp = o.bbpos; size_t t = 1;
    while (x < o.aepos) {
      int64_t y = std::min<int64_t>((x / L.tspace + 1) * L.tspace, o.aepos);
      int64_t blen = o.tr[t]; t += 2;
      WC.NP.align(a + x, (uint64_t)(y - x), b + bp, (uint64_t)blen);
      trace.insert(trace.end(), WC.NP.trace.begin(), WC.NP.trace.end());
      bp += blen; x = y;
    }
  }

  // top-D selection + ordering (src/daccord.cpp:2112-2288; conventions C2, C6)
  void selectOverlaps(uint64_t aread, std::vector<const OvlRec*>& sel) const {
    sel.clear();
    const Params& P = T.P;
    uint64_t b = L.first[aread], e = L.first[aread + 1];
    typedef std::pair<uint64_t, uint64_t> SI;
    struct Cmp { bool operator()(const SI& x, const SI& y) const { return x.first < y.first; } };
    FiniteHeap<SI, Cmp> H(P.maxinput);
    for (uint64_t i = b; i < e; ++i) {
      const OvlRec& o = L.ovl[i];
      uint64_t score = (uint64_t)std::ldexp((double)o.diffs / (double)(o.aepos - o.abpos), 30);
      if (H.f == P.maxinput) { if (score > H.top().first) H.popvoid(); }
      if (H.f < P.maxinput) H.push({score, i});
    }
    std::vector<uint64_t> idx;
    for (uint64_t i = 0; i < H.f; ++i) idx.push_back(H.H[i].second);
    std::sort(idx.begin(), idx.end());
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t x, uint64_t y) { return L.ovl[x].abpos < L.ovl[y].abpos; });
    for (auto i : idx) sel.push_back(&L.ovl[i]);
  }

  // HandleContext::operator() (src/HandleContext.hpp:1699-2901); appends FastA to out
  ReadStats handle(uint64_t aread, uint64_t& wellcounter, std::string& out) {
    ReadStats RS;
    const Params& P = T.P;
    std::vector<const OvlRec*> ita; selectOverlaps(aread, ita);
    const uint64_t nintv = ita.size();
    if (!nintv) return RS;
    const std::string& A = DB.reads[aread];
    uint64_t maxaepos = 0; double maxerate = 0.0, minerate = 1.0;
    for (auto o : ita) {
      if ((uint64_t)o->aepos > maxaepos) maxaepos = o->aepos;
      double er = (double)o->diffs / (double)(o->aepos - o->abpos);        // getErrorRate (inferred)
      if (er > maxerate) maxerate = er;
      if (er < minerate) minerate = er;
    }
    const double ediv = (maxerate > minerate) ? (maxerate - minerate) : 1.0;
    std::vector<std::vector<uint8_t>> traces(nintv);
    std::vector<std::string> bseq(nintv);
    std::map<uint64_t, ActiveElement> activeset;
    typedef std::pair<uint64_t, uint64_t> upair;
    FiniteHeap<upair> E(1024);
    std::vector<PileElement> PV;
    WindowsOf W(maxaepos, P.a, P.w);
    std::vector<SeqRef> MA;
    uint64_t z = 0;
    for (uint64_t y = 0; y < W.n; ++y) {
      const uint64_t astart = W.at(y).first, aend = W.at(y).second;
      while (z < nintv && (int64_t)astart >= ita[z]->abpos) {                       // :1904-1967
        const OvlRec& o = *ita[z];
        if (o.aepos >= (int64_t)astart) {
          bseq[z] = (o.flags & 1) ? revcomp(DB.reads[o.bread]) : DB.reads[o.bread];
          const uint8_t* ua0 = (const uint8_t*)A.data(); const uint8_t* ub0 = (const uint8_t*)bseq[z].data();
          computeTrace(o, ua0, ub0, traces[z]);
          const uint64_t aoff = astart - (uint64_t)o.abpos;
          const uint8_t* ta = traces[z].data(); const uint8_t* te = ta + traces[z].size();
          auto adv = advanceA(ta, te, aoff);
          uint64_t uboff = (uint64_t)o.bbpos + stringLengthUsed(ta, ta + adv.second).second;
          ta += adv.second;
          double er = (double)o.diffs / (double)(o.aepos - o.abpos);
          uint64_t escore = (uint64_t)(((er - minerate) / ediv) * std::numeric_limits<uint32_t>::max());
          uint64_t eindex = (escore << 32) | z;
          activeset[eindex] = ActiveElement{ua0 + astart, ub0 + uboff, ta, te, uboff};
          E.pushBump(upair((uint64_t)o.aepos, eindex));
        }
        ++z;
      }
      while (!E.empty() && E.top().first < aend) { upair UP = E.pop(); activeset.erase(UP.second); }   // :1969-1977
      MA.clear();
      const uint8_t* w_ua = activeset.size() ? activeset.begin()->second.ua : nullptr;
      for (auto& kv : activeset) {                                                 // :1984-2049
        ActiveElement& AE = kv.second;
        auto adv = advanceA(AE.ta, AE.te, P.w);
        uint64_t bwindowsize = stringLengthUsed(AE.ta, AE.ta + adv.second).second;
        auto advadv = advanceA(AE.ta, AE.te, W.offset(y));
        uint64_t badvancesize = stringLengthUsed(AE.ta, AE.ta + advadv.second).second;
        AE.ta += advadv.second;
        if (MA.empty()) MA.push_back(SeqRef(AE.ua, P.w));
        if (MA.size() < P.maxalign) MA.push_back(SeqRef(AE.ub, bwindowsize));
        AE.ua += W.offset(y); AE.ub += badvancesize; AE.uboff += badvancesize;
      }
      if (MA.empty()) continue;
      RS.windows++;
      WindowResult R = WC.run(MA.data(), MA.size());
      RS.attempted += R.attempted; RS.ok += R.ok;
      if (R.ok) {                                                                  // :2429-2493
        (void)w_ua;
        uint64_t apos = astart; size_t c = 0; const uint8_t* ta = R.trace.data(); const uint8_t* te = ta + R.trace.size();
        while (ta != te) {
          uint64_t numins = 0;
          while (ta != te && *ta == STEP_INS) { ++numins; ++ta; }
          for (uint64_t i = 0; i < numins; ++i) PV.push_back({(int64_t)apos, -(int64_t)numins + (int64_t)i, R.cons[c++]});
          if (ta != te) {
            switch (*(ta++)) {
              case STEP_MATCH: case STEP_MISMATCH: PV.push_back({(int64_t)apos++, 0, R.cons[c++]}); break;
              case STEP_DEL: PV.push_back({(int64_t)apos++, 0, 'D'}); break;
              default: break;
            }
          }
        }
      }
    }
    std::sort(PV.begin(), PV.end());                                               // :2541
    if (P.producefull) {                                                           // :2543-2580
      std::vector<PileElement> NPV; uint64_t next = 0; size_t low = 0;
      while (low < PV.size()) {
        size_t high = low + 1;
        while (high < PV.size() && PV[low].apos == PV[high].apos) ++high;
        for (; (int64_t)next < PV[low].apos; ++next) NPV.push_back({(int64_t)next, 0, (char)::tolower(A[next])});
        for (size_t i = low; i < high; ++i) NPV.push_back(PV[i]);
        next = PV[low].apos + 1; low = high;
      }
      for (; next < A.size(); ++next) NPV.push_back({(int64_t)next, 0, (char)::tolower(A[next])});
      PV.swap(NPV);
    }
    size_t il = 0;
    while (il < PV.size()) {                                                       // :2590-2612
      size_t ih = il + 1;
      while (ih != PV.size() && (PV[ih].apos - PV[ih - 1].apos) <= 1) ++ih;
      uint64_t first = PV[il].apos, last = PV[ih - 1].apos;
      if (last - first >= 100) {
        std::vector<char> CO; int64_t l = (int64_t)ih, depth = -1;
        while (l > (int64_t)il) {                                                  // :2627-2706
          int64_t h = --l;
          while (l >= 0 && PV[l].apos == PV[h].apos && PV[l].apre == PV[h].apre) --l;
          l += 1;
          uint64_t ld = (h - l) + 1;
          if (PV[l].apre == 0) depth = ld;
          std::pair<uint64_t, uint64_t> C[] = {{0, 'A'}, {0, 'C'}, {0, 'G'}, {0, 'T'}, {0, 'D'}, {0, 'a'}, {0, 'c'}, {0, 'g'}, {0, 't'}, {0, 0}};
          for (int64_t i = l; i <= h; ++i) {
            const char* q = strchr("ACGTDacgt", PV[i].sym);
            if (q && PV[i].sym) C[q - "ACGTDacgt"].first++;
          }
          for (int64_t i = ld; i < depth; ++i) C[4].first++;
          std::sort(&C[0], &C[10], std::greater<std::pair<uint64_t, uint64_t>>());
          if (C[0].first && C[0].second != 'D') CO.push_back((char)C[0].second);
        }
        std::reverse(CO.begin(), CO.end());
        if (P.producefull || CO.size() >= P.minlen) {                              // :2710-2724
          std::ostringstream o;
          o << '>' << (aread + 1) << '/' << wellcounter++ << '/' << first << '_' << first + CO.size() << " A=[" << first << "," << last << "]" << "\n";
          for (size_t zp = 0; zp < CO.size(); zp += 80) { o.write(CO.data() + zp, std::min<size_t>(80, CO.size() - zp)); o.put('\n'); }
          out += o.str();
        }
      }
      il = ih;
    }
    return RS;
  }
};

// ------------------------------------------------------------------ error profile estimation
// handleIndelEstimate<8> (reference src/daccord.cpp:271-631) and its driver (:1652-1880): windows of 40 advancing by 5 over
// the first <= 1024 A-reads, piles without a repeated 7-mer, consensus = the trivial unitig of the k=8 graph at frequency
// >= 2, every slice aligned to that consensus; the step counts are the profile.
struct AlignStats { uint64_t matches = 0, mismatches = 0, insertions = 0, deletions = 0;
  void add(const AlignStats& o) { matches += o.matches; mismatches += o.mismatches; insertions += o.insertions; deletions += o.deletions; }
  // libmaus2::lcs::AlignmentStatistics::getErrorRate (libmaus2 is not vendored; published definition)
  double errorRate() const { uint64_t e = mismatches + insertions + deletions; return (double)e / (double)(matches + e); } };
struct ProfileResult { AlignStats stats; uint64_t usable = 0, unusable = 0; double eavg = 0, edif = 0; uint64_t nreads = 0; };

// libmaus2::fastx::KmerRepeatDetector(k).detect (call site src/daccord.cpp:541): does any k-mer occur twice in the sequence (convention C9)
inline bool hasRepeatedKmer(const uint8_t* s, uint64_t n, unsigned k) {
  if (n < k) return false;
  std::vector<uint32_t> v;
  for (uint64_t i = 0; i + k <= n; ++i) { uint32_t w = 0; for (unsigned j = 0; j < k; ++j) w = (w << 2) | mapChar(s[i + j]); v.push_back(w); }
  std::sort(v.begin(), v.end());
  return std::adjacent_find(v.begin(), v.end()) != v.end();
}

struct ProfileEstimator {
  enum { EW = 40, EA = 5, EK = 8 };                 // src/daccord.cpp:1279-1280, :1587 (handleIndelEstimate<8>)
  const LasFile& L; const ReadDB& DB; uint64_t maxalign, maxinput;
  Aligner NP;
  KmerLimit KL; DebruijnGraph DG;
  ProfileEstimator(const LasFile& rL, const ReadDB& rDB, uint64_t rmaxalign, uint64_t rmaxinput)
      : L(rL), DB(rDB), maxalign(rmaxalign), maxinput(rmaxinput), KL(0.85, 0), DG(EK, 0.0, KL) {}

  // :1690-1741 -- keeps the D LOWEST scores (the comparator is the opposite of the main loop's), order by abpos (C6)
  void selectOverlaps(uint64_t aread, std::vector<const OvlRec*>& sel) const {
    sel.clear();
    typedef std::pair<uint64_t, uint64_t> SI;
    struct Cmp { bool operator()(const SI& x, const SI& y) const { return x.first > y.first; } };
    FiniteHeap<SI, Cmp> H(maxinput);
    std::vector<uint64_t> slot;
    for (uint64_t i = L.first[aread]; i < L.first[aread + 1]; ++i) {
      const OvlRec& o = L.ovl[i];
      uint64_t score = (uint64_t)std::ldexp((double)o.diffs / (double)(o.aepos - o.abpos), 30);
      if (H.f == maxinput) {
        if (score > H.top().first) continue;
        uint64_t p = H.top().second; H.popvoid(); slot[p] = i; H.push({score, p});
      } else { uint64_t p = slot.size(); H.push({score, p}); slot.push_back(i); }
    }
    std::sort(slot.begin(), slot.end());
    std::stable_sort(slot.begin(), slot.end(), [&](uint64_t x, uint64_t y) { return L.ovl[x].abpos < L.ovl[y].abpos; });
    for (auto i : slot) sel.push_back(&L.ovl[i]);
  }
  void computeTrace(const OvlRec& o, const uint8_t* a, const uint8_t* b, std::vector<uint8_t>& trace) {     // convention C8
    trace.clear();
    int64_t x = o.abpos, bp = o.bbpos; size_t t = 1;
    while (x < o.aepos) {
      int64_t y = std::min<int64_t>((x / L.tspace + 1) * L.tspace, o.aepos);
      int64_t blen = o.tr[t]; t += 2;
      NP.align(a + x, (uint64_t)(y - x), b + bp, (uint64_t)blen);
      trace.insert(trace.end(), NP.trace.begin(), NP.trace.end());
      bp += blen; x = y;
    }
  }

  // one A-read; returns the mean per-window error rate (0 if no window gave a consensus)
  double handleRead(uint64_t aread, ProfileResult& R) {
    std::vector<const OvlRec*> ita; selectOverlaps(aread, ita);
    const uint64_t nintv = ita.size();
    if (!nintv) return 0.0;
    const std::string& A = DB.reads[aread];
    double maxerate = 0.0, minerate = 1.0; uint64_t maxaepos = 0;
    std::vector<std::vector<uint8_t>> traces(nintv); std::vector<std::string> bseq(nintv);
    for (uint64_t z = 0; z < nintv; ++z) {
      const OvlRec& o = *ita[z];
      double er = (double)o.diffs / (double)(o.aepos - o.abpos);
      if (er > maxerate) maxerate = er;
      if (er < minerate) minerate = er;
      if ((uint64_t)o.aepos > maxaepos) maxaepos = (uint64_t)o.aepos;
      bseq[z] = (o.flags & 1) ? revcomp(DB.reads[o.bread]) : DB.reads[o.bread];
      computeTrace(o, (const uint8_t*)A.data(), (const uint8_t*)bseq[z].data(), traces[z]);          // all traces up front, :385-398
    }
    const double ediv = (maxerate > minerate) ? (maxerate - minerate) : 1.0;
    typedef std::pair<uint64_t, uint64_t> upair;
    FiniteHeap<upair> E(1024);
    std::map<uint64_t, ActiveElement> activeset;
    std::vector<SeqRef> MA;
    const uint64_t ylimit = (maxaepos + EA >= EW) ? ((maxaepos + EA - EW) / EA) : 0;                   // :408
    double esum = 0; uint64_t ecnt = 0, z = 0;
    for (uint64_t y = 0; y < ylimit; ++y) {
      const uint64_t astart = y * EA, aend = astart + EW;
      while (z < nintv && (int64_t)astart >= ita[z]->abpos) {                                       // :440-478
        const OvlRec& o = *ita[z];
        if (o.aepos >= (int64_t)astart) {
          const uint64_t aoff = astart - (uint64_t)o.abpos;
          const uint8_t* ta = traces[z].data(); const uint8_t* te = ta + traces[z].size();
          auto adv = advanceA(ta, te, aoff);
          uint64_t uboff = (uint64_t)o.bbpos + stringLengthUsed(ta, ta + adv.second).second;
          ta += adv.second;
          double er = (double)o.diffs / (double)(o.aepos - o.abpos);
          uint64_t escore = (uint64_t)(((er - minerate) / ediv) * std::numeric_limits<uint32_t>::max());
          uint64_t eindex = (escore << 32) | z;
          activeset[eindex] = ActiveElement{(const uint8_t*)A.data() + astart, (const uint8_t*)bseq[z].data() + uboff, ta, te, uboff};
          E.pushBump(upair((uint64_t)o.aepos, eindex));
        }
        ++z;
      }
      while (!E.empty() && E.top().first <= aend) { upair UP = E.pop(); activeset.erase(UP.second); }   // :481-485 (<=, unlike the main loop)
      MA.clear();
      for (auto& kv : activeset) {                                                                  // :490-530
        ActiveElement& AE = kv.second;
        auto adv = advanceA(AE.ta, AE.te, EW);
        uint64_t blen = stringLengthUsed(AE.ta, AE.ta + adv.second).second;
        if (MA.empty()) MA.push_back(SeqRef(AE.ua, EW));
        if (MA.size() < maxalign) MA.push_back(SeqRef(AE.ub, blen));
        auto advadv = advanceA(AE.ta, AE.te, EA);
        uint64_t badv = stringLengthUsed(AE.ta, AE.ta + advadv.second).second;
        AE.ua += EA; AE.ta += advadv.second; AE.ub += badv; AE.uboff += badv;
      }
      if (MA.size() < 3) continue;                                                                  // :532
      bool rep = false;
      for (auto& sr : MA) rep = hasRepeatedKmer(sr.first, sr.second, EK - 1) || rep;
      if (rep) { R.unusable++; continue; }
      R.usable++;
      DG.setup(MA.data(), MA.size());
      DG.filterFreq(2, MA.size());
      std::string cons;
      if (!DG.traverseTrivial(cons)) continue;
      AlignStats GAS;
      for (auto& sr : MA) {                                                                         // :585-600
        NP.align((const uint8_t*)cons.data(), cons.size(), sr.first, sr.second);
        for (uint8_t st : NP.trace) switch (st) {
          case STEP_MATCH: GAS.matches++; break; case STEP_MISMATCH: GAS.mismatches++; break;
          case STEP_INS: GAS.insertions++; break; default: GAS.deletions++; break; }
      }
      R.stats.add(GAS);
      esum += GAS.errorRate(); ecnt++;
    }
    return ecnt ? esum / ecnt : 0.0;
  }
};

// driver :1652-1880 over A-reads [lo, min(hi, lo+1024)) ; sequential accumulation order = read order
inline ProfileResult estimateProfile(const LasFile& L, const ReadDB& DB, int64_t lo, int64_t hi, uint64_t maxalign, uint64_t maxinput) {
  ProfileResult R; ProfileEstimator PE(L, DB, maxalign, maxinput);
  std::vector<double> loc;
  const int64_t top = std::min(hi, lo + 1024);
  for (int64_t r = lo; r < top; ++r) {
    if (L.first[r] == L.first[r + 1]) continue;
    double e = PE.handleRead((uint64_t)r, R); R.nreads++;
    if (e != 0.0) loc.push_back(e);
  }
  if (!loc.empty()) {
    double s = 0; for (double e : loc) s += e;
    R.eavg = s / loc.size();
    double d = 0; for (double e : loc) d += (R.eavg - e) * (R.eavg - e);
    R.edif = std::sqrt(d / loc.size());
  }
  return R;
}

}  // namespace oracle
