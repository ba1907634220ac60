// TEST INFRASTRUCTURE (oracle) -- parity unpinned, see oracle/README.md.
// One window of daccord's HandleContext::operator() (reference src/HandleContext.hpp):
//   expected length estimate  :2051-2155
//   k / filterfreq retry loop :2187-2351
//   consensus placement trace :2429-2493
#pragma once
#include <memory>
#include <map>
#include "dbg.hpp"

namespace oracle {

struct Params {                 // CLI surface, reference src/daccord.cpp:101-169, :1282-1305
  uint64_t w = 40, a = 10;
  uint64_t k_lo = 8, k_hi = 8;
  uint64_t maxalign = std::numeric_limits<uint64_t>::max();  // -d
  uint64_t minwindowcov = 3;                                  // -m
  uint64_t eminrate = std::numeric_limits<uint64_t>::max();   // -e
  uint64_t minlen = 0;                                        // -l
  int64_t minfilterfreq = 0, maxfilterfreq = 2;
  uint64_t maxinput = 5000;                                   // -D
  bool producefull = false;                                   // -f
  double p_i = 0.09, p_d = 0.045, est_cor = 0.85;             // from the error profile
};

struct WindowResult {
  bool attempted = false;       // MAo >= minwindowcov
  bool ok = false;              // !pathfailed
  int64_t elength = 0;
  uint64_t k = 0; int64_t filterfreq = -1; uint64_t minrate = 0;
  uint64_t ncand = 0;
  std::string cons;
  std::vector<uint8_t> trace;   // placement steps of align(Awindow, cons), forward order
};

// shared read-only tables + one graph per k (reference src/daccord.cpp:1867-1913, :1981-2023; ThreadContext HandleContext.hpp:276-319)
struct Tables {
  Params P;
  OffsetLikely OL;
  std::map<uint64_t, KmerLimit> MKL;
  explicit Tables(const Params& rP) : P(rP), OL(computeOffsetLikely(rP.w, rP.p_i, rP.p_d)) {
    for (uint64_t k = P.k_lo; k <= P.k_hi; ++k) MKL.emplace(k, KmerLimit(std::pow(P.est_cor, (double)k), 100));
  }
};

struct WindowContext {
  const Tables& T;
  std::vector<std::unique_ptr<DebruijnGraph>> ADG;
  Aligner NP;
  explicit WindowContext(const Tables& rT) : T(rT) {
    for (uint64_t k = T.P.k_lo; k <= T.P.k_hi; ++k) ADG.emplace_back(new DebruijnGraph((unsigned)k, T.P.est_cor, T.MKL.at(k)));
  }

  // :2051-2155
  int64_t estimateLength(const SeqRef* MA, uint64_t MAo) const {
    const OffsetLikely& OL = T.OL;
    int64_t maxvprodindex = -1;
    if (MAo) {
      int64_t minSup = (int64_t)MA[0].second - 1, maxSup = minSup;
      for (uint64_t j = 1; j < MAo; ++j) { int64_t lp = (int64_t)MA[j].second - 1; minSup = std::min(minSup, lp); maxSup = std::max(maxSup, lp); }
      if (minSup < 0) minSup = 0;
      if (maxSup < 0) maxSup = 0;
      uint64_t supStart = OL.supportLow(minSup), supEnd = OL.supportHigh(maxSup);
      double maxval = std::numeric_limits<double>::min();
      for (uint64_t i = supStart; i < supEnd; ++i) {
        const DotVec& DP = OL.DPnorm[i];
        double vprod = 1.0;
        for (uint64_t j = 0; j < MAo; ++j) { uint64_t len = MA[j].second; if (len) vprod *= DP.at(len - 1); }
        if (vprod > maxval) { maxval = vprod; maxvprodindex = (int64_t)i; }
      }
    }
    if (maxvprodindex == -1) {                       // :2103-2155
      int64_t maxoff = -1; double maxoffv = std::numeric_limits<double>::min();
      std::vector<uint64_t> Vdist;
      for (uint64_t i = 0; i < MAo; ++i) Vdist.push_back(MA[i].second);
      std::sort(Vdist.begin(), Vdist.end());
      std::vector<double> VVVV;
      uint64_t low = 0;
      while (low < Vdist.size()) {
        uint64_t high = low + 1;
        while (high < Vdist.size() && Vdist[high] == Vdist[low]) ++high;
        while (!(Vdist[low] < VVVV.size())) VVVV.push_back(0);
        VVVV[Vdist[low]] = (double)((high - low) - 1);
        low = high;
      }
      for (uint64_t i = 0; i < OL.DPnormSquare.size(); ++i) {
        double v = OL.DPnormSquare[i].dot(VVVV.data(), VVVV.size());
        if (v > maxoffv) { maxoff = (int64_t)i; maxoffv = v; }
      }
      if (maxoff != -1 && maxoffv >= 1e-3) maxvprodindex = maxoff;
    }
    return maxvprodindex + 1;
  }

  // the per-window body of HandleContext::operator() :2164-2494 ; MA[0] is the A window (HandleContext.hpp:2032-2043)
  WindowResult run(const SeqRef* MA, uint64_t MAo) {
    const Params& P = T.P;
    WindowResult R;
    R.elength = estimateLength(MA, MAo);
    if (!(MAo >= P.minwindowcov)) return R;
    R.attempted = true;
    bool pathfailed = true;
    uint64_t minindex = 0, minrate = P.eminrate;
    DebruijnGraph* minDG = nullptr;
    int64_t usedff = -1;
    for (auto& pDG : ADG) {
      DebruijnGraph& DG = *pDG;
      for (int64_t filterfreq = P.maxfilterfreq; filterfreq >= P.minfilterfreq; --filterfreq) {
        DG.setup(MA, MAo);
        DG.filterFreq((uint64_t)std::max(filterfreq, (int64_t)1), MAo);
        DG.computeFeasibleKmerPositions(T.OL, 1e-3);
        if (filterfreq == 0) {
          DG.getLevelSuccessors(2);
          DG.setupNodes();
          DG.setupAddHeap(MAo);
          DG.computeFeasibleKmerPositions(T.OL, 1e-3);
        }
        uint64_t mintry = 0; const uint64_t maxtries = 3; bool lconsok = false;
        do {
          bool consok = DG.traverse(R.elength - 4, R.elength + 4, MA, MAo, 16, 16);
          if (consok) {
            auto MR = DG.checkCandidatesU(MA, MAo);
            if (MR.second < minrate) { lconsok = true; minrate = MR.second; minindex = MR.first; minDG = &DG; usedff = filterfreq; }
            else if (minDG) lconsok = true;
            break;
          } else if (++mintry >= maxtries) break;
        } while (DG.addNextFromHeap());
        if (lconsok) { pathfailed = false; break; }
      }
    }
    if (!pathfailed) {
      auto c = minDG->getCandidate(minindex);
      R.ok = true; R.k = minDG->getKmerSize(); R.filterfreq = usedff; R.minrate = minrate; R.ncand = minDG->getNumCandidates();
      R.cons.assign((const char*)c.first, (const char*)c.second);
      NP.align(MA[0].first, P.w, c.first, (uint64_t)(c.second - c.first));     // :2434
      R.trace = NP.trace;
    }
    return R;
  }
};

}  // namespace oracle
