// vote.hpp -- consensus placement records, pile vote and FastA assembly for one A-read
// (reference src/HandleContext.hpp:2446-2493 placement walk, :2541-2724 vote and output).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <algorithm>
#include <cctype>
#include "../../../include/daccord_b200.h"

namespace dhost {

struct PileElement {                       // HandleContext::PileElement (src/HandleContext.hpp:219-248)
  int64_t apos, apre; char sym;
  bool operator<(const PileElement& o) const { return apos != o.apos ? apos < o.apos : (apre != o.apre ? apre < o.apre : sym < o.sym); }
};

// walk the placement trace of one window (:2446-2493)
inline void place_window(const dcu_window& W, const dcu_result& R, const uint8_t* cons, const uint8_t* ops, std::vector<PileElement>& PV) {
  int64_t apos = W.astart; uint32_t c = 0, t = 0;
  while (t < R.nops) {
    uint32_t numins = 0;
    while (t < R.nops && ops[t] == DCU_STEP_INS) { ++numins; ++t; }
    for (uint32_t i = 0; i < numins; ++i) PV.push_back({apos, -(int64_t)numins + (int64_t)i, (char)cons[c++]});
    if (t < R.nops) {
      uint8_t op = ops[t++];
      if (op == DCU_STEP_MATCH || op == DCU_STEP_MISMATCH) PV.push_back({apos++, 0, (char)cons[c++]});
      else if (op == DCU_STEP_DEL) PV.push_back({apos++, 0, 'D'});
    }
  }
}

struct VoteParams { bool producefull = false; uint64_t minlen = 0; };

// sort, (optionally) fill uncorrected stretches in lower case, split into runs, majority per column, FastA
// aread_bases: the A read as ASCII (needed for -f only); counter: the shared wellcounter (:2712)
inline void vote_read(int64_t aid, std::vector<PileElement>& PV, const VoteParams& VP, const std::string& aread_bases, uint64_t& counter, std::string& out) {
  std::sort(PV.begin(), PV.end());
  if (VP.producefull) {                                          // :2543-2580
    std::vector<PileElement> NPV; uint64_t next = 0; size_t low = 0;
    while (low < PV.size()) {
      size_t high = low + 1;
      while (high < PV.size() && PV[low].apos == PV[high].apos) ++high;
      for (; (int64_t)next < PV[low].apos; ++next) NPV.push_back({(int64_t)next, 0, (char)::tolower(aread_bases[next])});
      for (size_t i = low; i < high; ++i) NPV.push_back(PV[i]);
      next = (uint64_t)PV[low].apos + 1; low = high;
    }
    for (; next < aread_bases.size(); ++next) NPV.push_back({(int64_t)next, 0, (char)::tolower(aread_bases[next])});
    PV.swap(NPV);
  }
  size_t il = 0;
  while (il < PV.size()) {                                        // runs :2590-2612
    size_t ih = il + 1;
    while (ih != PV.size() && (PV[ih].apos - PV[ih - 1].apos) <= 1) ++ih;
    const uint64_t first = (uint64_t)PV[il].apos, last = (uint64_t)PV[ih - 1].apos;
    if (last - first >= 100) {
      std::string CO;
      int64_t l = (int64_t)ih, depth = -1;
      while (l > (int64_t)il) {                                   // :2627-2706, right to left
        int64_t h = --l;
        while (l >= 0 && PV[l].apos == PV[h].apos && PV[l].apre == PV[h].apre) --l;
        l += 1;
        const int64_t ld = (h - l) + 1;
        if (PV[l].apre == 0) depth = ld;
        std::pair<uint64_t, uint64_t> C[10] = {{0, 'A'}, {0, 'C'}, {0, 'G'}, {0, 'T'}, {0, 'D'}, {0, 'a'}, {0, 'c'}, {0, 'g'}, {0, 't'}, {0, 0}};
        for (int64_t i = l; i <= h; ++i) switch (PV[i].sym) {
          case 'A': C[0].first++; break; case 'C': C[1].first++; break; case 'G': C[2].first++; break; case 'T': C[3].first++; break;
          case 'D': C[4].first++; break; case 'a': C[5].first++; break; case 'c': C[6].first++; break; case 'g': C[7].first++; break;
          case 't': C[8].first++; break; default: break;
        }
        for (int64_t i = ld; i < depth; ++i) C[4].first++;
        std::sort(&C[0], &C[10], std::greater<std::pair<uint64_t, uint64_t>>());
        if (C[0].first && C[0].second != 'D') CO.push_back((char)C[0].second);
      }
      std::reverse(CO.begin(), CO.end());
      if (VP.producefull || CO.size() >= VP.minlen) {             // :2710-2724
        out += ">" + std::to_string(aid + 1) + "/" + std::to_string(counter++) + "/" + std::to_string(first) + "_" + std::to_string(first + CO.size()) +
               " A=[" + std::to_string(first) + "," + std::to_string(last) + "]\n";
        for (size_t z = 0; z < CO.size(); z += 80) { out.append(CO, z, std::min<size_t>(80, CO.size() - z)); out.push_back('\n'); }
      }
    }
    il = ih;
  }
}

// FastA text of the segments the GPU vote returns (dcu_vote / dcu_get_corrected): same header and line layout as vote_read (:2710-2724)
inline void format_segments(const dcu_segment* seg, uint64_t nseg, const char* chars, uint64_t& counter, std::string& out) {
  for (uint64_t i = 0; i < nseg; ++i) {
    const dcu_segment& g = seg[i];
    out += ">" + std::to_string((uint64_t)g.aread + 1) + "/" + std::to_string(counter++) + "/" + std::to_string(g.first) + "_" + std::to_string((uint64_t)g.first + g.len) +
           " A=[" + std::to_string(g.first) + "," + std::to_string(g.last) + "]\n";
    for (uint64_t z = 0; z < g.len; z += 80) { out.append(chars + g.off + z, (size_t)std::min<uint64_t>(80, g.len - z)); out.push_back('\n'); }
  }
}

}  // namespace dhost
