// TEST INFRASTRUCTURE (oracle) -- parity unpinned, see oracle/README.md.
// CPU restatement of daccord's per-window local de Bruijn graph (k is a run-time value here;
// the reference instantiates DebruijnGraph<k> for k in [3,12], src/DebruijnGraphContainer.hpp:41-110).
// Every method cites the reference lines it follows (all in src/DebruijnGraph.hpp unless noted).
#pragma once
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <cassert>
#include <vector>
#include <string>
#include <limits>
#include <algorithm>
#include <numeric>
#include <stdexcept>
#include "tables.hpp"
#include "heap.hpp"
#include "align.hpp"

namespace oracle {

typedef std::pair<const uint8_t*, uint64_t> SeqRef;   // (ASCII bases, length)

inline unsigned mapChar(uint8_t c) {   // libmaus2::fastx::mapChar (A,C,G,T -> 0..3)
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 0; }
}
inline uint8_t remapChar(unsigned v) { return "ACGT"[v & 3]; }

// src/Node.hpp:21-58
struct Node {
  uint64_t v = 0, spo = 0, freq = 0, numsucc = 0, numsuccactive = 0;
  uint64_t feaspos = 0, cfeaspos = 0, numfeaspos = 0, numcfeaspos = 0;
  uint64_t pfostart = 0, pfosize = 0, cpfostart = 0, cpfosize = 0;
  uint64_t plow = 0, phigh = 0, cplow = 0, cphigh = 0;
};

// src/Links.hpp:23-73 : <=4 successors packed (freq<<8)|sym, sorted descending
struct Links {
  uint64_t A[4]; uint64_t p = 0;
  void reset() { p = 0; }
  void push(uint64_t sym, uint64_t freq) { if (freq) A[p++] = (freq << 8) | sym; }
  void sort() { if (p > 1) std::sort(A, A + p, std::greater<uint64_t>()); }
  uint64_t size() const { return p; }
  uint64_t getFreq(uint64_t i) const { return A[i] >> 8; }
  uint64_t getSym(uint64_t i) const { return A[i] & 0xFF; }
};

struct Stretch {                                   // :85-185
  uint64_t first = 0, ext = 0, last = 0, len = 0, stretchO = 0;
  uint64_t feasposO = 0, feasposL = 0, cfeasposO = 0, cfeasposL = 0;
  bool operator<(const Stretch& O) const {
    if (first != O.first) return first < O.first;
    if (ext != O.ext) return ext < O.ext;
    if (len != O.len) return len > O.len;
    return last < O.last;
  }
  bool operator==(const Stretch& O) const { return first == O.first && ext == O.ext && last == O.last && len == O.len; }
};
struct Path { uint64_t len = 0, off = 0, pos = 0; double weight = 0.0; uint64_t baselen = 0; };     // :204-240
struct ReversePath {                               // :261-324
  uint32_t linkoff = 0, front = 0; double weight = 0.0; uint16_t pos = 0, len = 0, baselen = 0;
};
struct EdgeActivationElement {                     // :403-423
  uint64_t freq, nodeid, edgeid;
  bool operator<(const EdgeActivationElement& E) const {
    if (E.freq != freq) return freq > E.freq;
    if (nodeid != E.nodeid) return nodeid < E.nodeid;
    return edgeid < E.edgeid;
  }
};
struct ConsensusCandidate { uint64_t o = 0, l = 0; double weight = 0, error = 0; };                 // :434-458
struct ScoreInterval { uint64_t left, right, current; double weight; Path P; };                     // :495-518
struct SeqPos { uint32_t seq, pos; bool operator<(const SeqPos& O) const { return pos != O.pos ? pos < O.pos : seq < O.seq; } };
struct PosFreq { uint32_t pos, freq; };
struct LevelAddElement { uint64_t from, to, v, off; };
struct NodeAddElement { uint64_t v, pos; bool operator<(const NodeAddElement& O) const { return v != O.v ? v < O.v : pos < O.pos; } };
struct StretchFeasObject { uint64_t p; double w, wf, wl; };                                         // :875-889

struct CmpPathWeightLess { bool operator()(const Path& A, const Path& B) const { return A.weight < B.weight; } };          // :253-259
struct CmpRPWeightLess { bool operator()(const ReversePath& A, const ReversePath& B) const { return A.weight < B.weight; } };    // :350-356
struct CmpRPWeightGreater { bool operator()(const ReversePath& A, const ReversePath& B) const { return A.weight > B.weight; } }; // :366-372
struct CmpSIWeightGreater { bool operator()(const ScoreInterval& A, const ScoreInterval& B) const { return A.weight > B.weight; } }; // :514-517
struct CmpCCWeightLess { bool operator()(const ConsensusCandidate& A, const ConsensusCandidate& B) const { return A.weight < B.weight; } };    // :461-467
struct CmpCCWeightGreater { bool operator()(const ConsensusCandidate& A, const ConsensusCandidate& B) const { return A.weight > B.weight; } }; // :469-475

enum { CDH_SIZE = 16, CD_SIZE = 16, REVERSE_PATH_HEAP_SIZE = 12, PATH_HEAP_SIZE = 12 };             // :2333-2336

class DebruijnGraph {
 public:
  const unsigned kmersize;
  const double p;            // est_cor
  KmerLimit KL;
  const uint64_t m;          // k-mer mask

  std::vector<uint64_t> prenodes, last;
  std::vector<uint32_t> seqlen;
  uint64_t maxk = 0;
  std::vector<SeqPos> SP, RSP;
  std::vector<PosFreq> PF, RPF;
  std::vector<Node> nodes;
  std::vector<int32_t> nodecache;
  FiniteHeap<EdgeActivationElement> EAH{1024};
  std::vector<uint64_t> stretchLinks;
  std::vector<Stretch> stretches;
  std::vector<uint8_t> Acons;
  std::vector<uint64_t> AP, APR;
  std::vector<ConsensusCandidate> ACC;
  std::vector<std::pair<uint64_t, double>> Afeaspos, Acfeaspos;
  std::vector<StretchFeasObject> Astretchfeas, Acstretchfeas;
  std::vector<std::pair<uint64_t, uint64_t>> reverseStretchLinks;
  std::vector<ReversePath> ARP;
  std::vector<uint64_t> ARW;
  uint64_t maxkmerpos = 0, maxstretchlength = 0, maxsupto = 0;
  std::vector<FiniteHeap<ReversePath, CmpRPWeightLess>> ARPH;
  std::vector<FiniteHeap<Path, CmpPathWeightLess>> APQ;
  FiniteHeap<ReversePath, CmpRPWeightGreater> RPST{1024};
  FiniteHeap<ScoreInterval, CmpSIWeightGreater> SIQ{1024};
  FiniteHeap<ConsensusCandidate, CmpCCWeightLess> CDH{CDH_SIZE};
  FiniteHeap<ConsensusCandidate, CmpCCWeightGreater> CH{CD_SIZE};
  std::vector<std::pair<uint64_t, uint64_t>> maxFirst, maxLast;

  DebruijnGraph(unsigned k, double est_cor, const KmerLimit& rKL)      // :2359-2383
      : kmersize(k), p(est_cor), KL(rKL), m((k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1)) {
    if (k < 1 || k > 14) throw std::runtime_error("oracle: k out of range");
    nodecache.assign(1ull << (2 * k), -1);
  }

  uint64_t getKmerSize() const { return kmersize; }
  static uint64_t combine(uint64_t v, uint64_t seq, uint64_t pos) { return (v << 32) | (pos << 16) | seq; }   // :1209-1217
  static uint64_t kmerMask(uint64_t w) { return w >> 32; }
  static uint64_t seqMask(uint64_t w) { return w & 0xFFFF; }
  static uint64_t posMask(uint64_t w) { return (w >> 16) & 0xFFFF; }

  const Node* getNode(uint64_t v) const { int32_t j = nodecache[v]; return j < 0 ? nullptr : &nodes[j]; }   // :968-985
  int64_t getNodeId(uint64_t v) const { return nodecache[v]; }
  uint64_t count(uint64_t v) const { const Node* n = getNode(v); return n ? n->freq : 0; }                     // :2388-2396
  void clearNodeCache() { for (auto& n : nodes) nodecache[n.v] = -1; }                                        // :1163-1172
  void setupNodeCache() { for (size_t i = 0; i < nodes.size(); ++i) nodecache[nodes[i].v] = (int32_t)i; }     // :1174-1178

  // ---- build: :2018-2331 ----
  void setupPreNodes(const SeqRef* I, uint64_t o) {          // :2018-2304 (the LSD radix sort there == ascending sort of the words)
    prenodes.clear(); last.clear(); seqlen.clear(); maxk = 0;
    for (uint64_t j = 0; j < o; ++j) {
      if (I[j].second >= kmersize) {
        uint64_t numk = I[j].second - kmersize + 1;
        const uint8_t* u = I[j].first;
        uint64_t v = 0;
        for (unsigned i = 0; i + 1 < kmersize; ++i) { v <<= 2; v |= mapChar(*(u++)); }
        for (uint64_t i = 0; i < numk; ++i) {
          v <<= 2; v &= m; v |= mapChar(*(u++));
          prenodes.push_back(combine(v, j, i));
        }
        last.push_back(combine(v, j, numk - 1));
        maxk = std::max(maxk, numk);
      }
      seqlen.push_back((uint32_t)I[j].second);
    }
    std::sort(prenodes.begin(), prenodes.end());
    std::sort(last.begin(), last.end());
  }

  void setupNodes() {                                          // :1918-2014
    clearNodeCache();
    nodes.clear(); SP.clear(); RSP.clear(); PF.clear(); RPF.clear();
    maxkmerpos = 0;
    uint64_t l = 0, np = prenodes.size();
    while (l < np) {
      uint64_t h = l, li = l, lp = posMask(prenodes[l]);
      uint64_t pfostart = PF.size(), rpfostart = RPF.size();
      while (h < np && kmerMask(prenodes[h]) == kmerMask(prenodes[l])) {
        uint64_t seq = seqMask(prenodes[h]), pos = posMask(prenodes[h]);
        if (pos != lp) { PF.push_back({(uint32_t)lp, (uint32_t)(h - li)}); lp = pos; li = h; }
        SP.push_back({(uint32_t)seq, (uint32_t)pos});
        RSP.push_back({(uint32_t)seq, (uint32_t)(seqlen[seq] - pos - kmersize)});
        ++h;
      }
      PF.push_back({(uint32_t)lp, (uint32_t)(h - li)});
      uint64_t freq = h - l;
      std::sort(RSP.end() - freq, RSP.end());
      uint64_t cl = RSP.size() - freq;
      while (cl < RSP.size()) {
        uint64_t ch = cl + 1;
        while (ch < RSP.size() && RSP[ch].pos == RSP[cl].pos) ++ch;
        RPF.push_back({RSP[cl].pos, (uint32_t)(ch - cl)});
        cl = ch;
      }
      Node node;
      node.v = kmerMask(prenodes[l]); node.spo = SP.size() - freq; node.freq = freq;
      node.pfostart = pfostart; node.pfosize = PF.size() - pfostart;
      node.cpfostart = rpfostart; node.cpfosize = RPF.size() - rpfostart;
      node.plow = PF[pfostart].pos; node.phigh = PF.back().pos;
      node.cplow = RPF[rpfostart].pos; node.cphigh = RPF.back().pos;
      nodes.push_back(node);
      maxkmerpos = std::max(maxkmerpos, std::max(node.cphigh, node.phigh));
      l = h;
    }
    setupNodeCache();
  }

  void getSuccessors(uint64_t v, Links& L) const {              // :2413-2426
    L.reset();
    uint64_t masked = (v << 2) & m;
    for (uint64_t i = 0; i < 4; ++i) L.push(i, count(masked | i));
    L.sort();
  }
  void getActiveSuccessors(uint64_t v, Links& L) const {        // :2434-2460
    L.reset();
    const Node* node = getNode(v);
    if (node) { getSuccessors(v, L); L.p = node->numsuccactive; } else L.p = 0;
  }
  uint64_t getUniqueActiveSuccessor(uint64_t v) const { Links L; getActiveSuccessors(v, L); return ((v << 2) & m) | L.getSym(0); }  // :2468-2476
  uint64_t getNumActiveSuccessors(uint64_t v) const { Links L; getActiveSuccessors(v, L); return L.size(); }                          // :2479-2484
  bool isEdgeActive(uint64_t from, uint64_t to) const {         // :2487-2503
    Links L; getActiveSuccessors(from, L);
    uint64_t masked = (from << 2) & m;
    for (uint64_t i = 0; i < L.size(); ++i) if (to == (masked | L.getSym(i))) return true;
    return false;
  }
  void getPredecessors(uint64_t v, Links& L) const {            // :2535-2550
    L.reset();
    uint64_t masked = (v >> 2) & m; unsigned shift = 2 * (kmersize - 1);
    for (uint64_t i = 0; i < 4; ++i) L.push(i, count(masked | (i << shift)));
    L.sort();
  }
  uint64_t getNumActivePredecessors(uint64_t v) const {         // :2552-2597
    Links L; L.reset();
    if (!getNode(v)) return 0;
    uint64_t masked = (v >> 2) & m; unsigned shift = 2 * (kmersize - 1);
    getPredecessors(v, L);
    uint64_t o = 0;
    for (uint64_t i = 0; i < L.size(); ++i) if (isEdgeActive(masked | (L.getSym(i) << shift), v)) ++o;
    return o;
  }

  void setNodesActive(bool check, uint64_t lim) {               // :1770-1814
    Links L;
    for (auto& node : nodes) {
      getSuccessors(node.v, L);
      if (L.size()) {
        node.numsucc = L.size(); node.numsuccactive = 1;
        while (node.numsuccactive < L.size() &&
               ((L.getFreq(node.numsuccactive) >= L.getFreq(0) / 2) || (check && (L.getFreq(node.numsuccactive) >= lim))))
          ++node.numsuccactive;
      } else { node.numsucc = 0; node.numsuccactive = 0; }
    }
  }
  void setupAddHeap(uint64_t no) {                              // :1818-1859
    for (auto& node : nodes) { node.numsucc = 0; node.numsuccactive = 0; }
    if (p) setNodesActive(true, (uint64_t)KL.getLimit(no)); else setNodesActive(false, 0);
    EAH.clear();
    Links L;
    for (uint64_t z = 0; z < nodes.size(); ++z) {
      getSuccessors(nodes[z].v, L);
      for (uint64_t i = nodes[z].numsuccactive; i < L.size(); ++i) EAH.pushBump({L.getFreq(i), z, i});
    }
  }
  bool addNextFromHeap() {                                      // :1861-1897
    if (EAH.empty()) return false;
    uint64_t topfreq = EAH.top().freq;
    while (!EAH.empty() && EAH.top().freq == topfreq) { EdgeActivationElement E = EAH.pop(); nodes[E.nodeid].numsuccactive += 1; }
    return true;
  }
  void setup(const SeqRef* I, uint64_t o) {                     // :2307-2331
    stretches.clear(); clearNodeCache(); nodes.clear(); EAH.clear();
    setupPreNodes(I, o); setupNodes(); setupAddHeap(o);
  }
  void filterFreq(uint64_t f, uint64_t no) {                    // :1181-1197
    clearNodeCache();
    uint64_t o = 0;
    for (uint64_t i = 0; i < nodes.size(); ++i) if (nodes[i].freq >= f) nodes[o++] = nodes[i];
    nodes.resize(o);
    setupNodeCache();
    setupAddHeap(no);
  }

  // ---- positional feasibility: :3117-3174, :3826-3904 ----
  double kmerPositionWeight(const Node& node, uint64_t pp, const OffsetLikely& OL, bool reverse) const {
    if (pp >= OL.size()) return 0;
    const DotVec& DP = OL.DPnormSquare[pp];
    const PosFreq* q = (reverse ? RPF.data() + node.cpfostart : PF.data() + node.pfostart);
    const PosFreq* qe = q + (reverse ? node.cpfosize : node.pfosize);
    while (q != qe && q->pos < DP.firstsign) ++q;
    uint64_t e = DP.firstsign + DP.V.size();
    uint64_t uprr = 0;
    for (; q != qe && q->pos < e; ++q) uprr += (uint64_t)q->freq * DP.VS[q->pos - DP.firstsign];
    return (double)uprr / 4294967296.0;
  }
  void computeFeasibleKmerPositions(const OffsetLikely& OL, double thres) {
    Afeaspos.clear(); Acfeaspos.clear(); maxsupto = 0;
    for (auto& node : nodes) {
      node.feaspos = Afeaspos.size(); node.cfeaspos = Acfeaspos.size();
      uint64_t pfrom = OL.supportLow(node.plow), pto = OL.supportHigh(node.phigh);
      maxsupto = std::max(maxsupto, pto);
      for (uint64_t pp = pfrom; pp < pto; ++pp) {
        double w = kmerPositionWeight(node, pp, OL, false);
        if (w >= thres) Afeaspos.push_back({pp, w});
      }
      uint64_t cpfrom = OL.supportLow(node.cplow), cpto = OL.supportHigh(node.cphigh);
      maxsupto = std::max(maxsupto, cpto);
      for (uint64_t pp = cpfrom; pp < cpto; ++pp) {
        double w = kmerPositionWeight(node, pp, OL, true);
        if (w >= thres) Acfeaspos.push_back({pp, w});
      }
      node.numfeaspos = Afeaspos.size() - node.feaspos;
      node.numcfeaspos = Acfeaspos.size() - node.cfeaspos;
    }
  }

  // ---- gap filling: :1016-1161 ----
  void getLevelSuccessors(unsigned s) {
    std::vector<LevelAddElement> LS;
    for (uint64_t i = 0; i < nodes.size(); ++i) {
      uint64_t v = nodes[i].v;
      uint64_t low = (v << (2 * s)) & m, high = low | ((1ull << (2 * s)) - 1);
      auto nlow = std::lower_bound(nodes.begin(), nodes.end(), low, [](const Node& a, uint64_t b) { return a.v < b; });
      auto nhigh = std::upper_bound(nodes.begin(), nodes.end(), high, [](uint64_t a, const Node& b) { return a < b.v; });
      for (auto np = nlow; np != nhigh; ++np) {
        uint64_t nv = np->v;
        for (unsigned ii = 1; ii < s; ++ii) {
          uint64_t vhigh = (v << (2 * ii)) & m, vlow = nv >> ((s - ii) * 2);
          uint64_t cv = vlow | vhigh;
          if (!getNode(cv)) LS.push_back({v, nv, cv, ii});
        }
      }
    }
    std::vector<NodeAddElement> ANE;
    std::vector<std::pair<uint64_t, double>> T;
    for (auto& L : LS) {
      const Node& from = *getNode(L.from); const Node& to = *getNode(L.to);
      T.clear();
      for (uint64_t j = 0; j < from.numfeaspos; ++j) T.push_back({Afeaspos[from.feaspos + j].first + s, Afeaspos[from.feaspos + j].second});
      for (uint64_t j = 0; j < to.numfeaspos; ++j) T.push_back(Afeaspos[to.feaspos + j]);
      std::sort(T.begin(), T.end());
      uint64_t l = 0, mp = 0; double mweight = std::numeric_limits<double>::min();
      while (l < T.size()) {
        uint64_t h = l + 1;
        while (h < T.size() && T[l].first == T[h].first) ++h;
        if (h - l > 1) {
          uint64_t pp = T[l].first;
          uint64_t pos = pp - s + L.off;
          double weight = T[l].second + T[h - 1].second;
          if (weight > mweight) { mweight = weight; mp = pos; }
        }
        l = h;
      }
      if (mweight != std::numeric_limits<double>::min()) ANE.push_back({L.v, mp});
    }
    std::sort(ANE.begin(), ANE.end());
    for (auto& a : ANE) {
      int64_t seqid = -1;
      for (uint64_t j = 0; j < seqlen.size() && seqid < 0; ++j) if (a.pos + kmersize <= seqlen[j]) seqid = (int64_t)j;
      if (seqid != -1) prenodes.push_back(combine(a.v, (uint64_t)seqid, a.pos));
    }
    std::sort(prenodes.begin(), prenodes.end());
  }

  // ---- first / last k-mer lists: :1280-1304, :1360-1391 ----
  void maxForPosList(uint64_t pos) {
    maxFirst.clear();
    for (auto& node : nodes) {
      uint64_t c = 0;
      for (uint64_t j = 0; j < node.freq; ++j) if (SP[node.spo + j].pos == pos) ++c;
      if (c) maxFirst.push_back({c, node.v});
    }
    std::sort(maxFirst.begin(), maxFirst.end(), std::greater<std::pair<uint64_t, uint64_t>>());
  }
  void maxLastList() {
    maxLast.clear();
    uint64_t l = 0;
    while (l < last.size()) {
      uint64_t h = l + 1;
      while (h < last.size() && (last[h] >> 32) == (last[l] >> 32)) ++h;
      maxLast.push_back({h - l, last[l] >> 32});
      l = h;
    }
    std::sort(maxLast.begin(), maxLast.end(), std::greater<std::pair<uint64_t, uint64_t>>());
  }

  // ---- stretches: :2599-2637, :2772-2841, :2844-2986, :3087-3114 ----
  void copyStretch(uint64_t low, uint64_t high) {
    Stretch s; s.first = stretchLinks[low]; s.ext = stretchLinks[low + 1]; s.last = stretchLinks[high - 1]; s.len = high - low;
    s.stretchO = stretchLinks.size();
    for (uint64_t i = low; i < high; ++i) { uint64_t link = stretchLinks[i]; stretchLinks.push_back(link); }
    stretches.push_back(s);
  }
  void splitStretches(uint64_t v) {
    std::vector<uint64_t> splitA;
    uint64_t loopend = stretches.size();
    for (uint64_t z = 0; z < loopend; ++z) {
      Stretch st = stretches[z];
      int64_t splitindex = -1;
      for (uint64_t i = 1; i + 1 < st.len; ++i) if (stretchLinks[st.stretchO + i] == v) { splitindex = (int64_t)i; break; }
      if (splitindex != -1) {
        copyStretch(st.stretchO, st.stretchO + splitindex + 1);
        copyStretch(st.stretchO + splitindex, st.stretchO + st.len);
        splitA.push_back(z);
      }
    }
    uint64_t l = 0, idx = 0, o = 0;
    for (; idx < splitA.size(); ++l) { if (l == splitA[idx]) ++idx; else stretches[o++] = stretches[l]; }
    while (l < stretches.size()) stretches[o++] = stretches[l++];
    stretches.resize(o);
  }
  void computeStretches(bool checkpredecessors) {
    std::vector<uint8_t> BV(nodes.size(), 0);
    stretchLinks.clear(); stretches.clear(); maxstretchlength = 0;
    for (uint64_t z = 0; z < nodes.size(); ++z) {
      const uint64_t refk = nodes[z].v;
      uint64_t numpred = getNumActivePredecessors(refk);
      uint64_t numsucc = nodes[z].numsuccactive;
      if (numsucc && (numpred != 1 || numsucc > 1)) {
        Links L; getActiveSuccessors(refk, L);
        for (uint64_t i = 0; i < numsucc; ++i) {
          uint64_t start = stretchLinks.size();
          uint64_t firstext = ((refk << 2) & m) | L.getSym(i);
          uint64_t extk = firstext;
          stretchLinks.push_back(refk); BV[getNodeId(refk)] = 1;
          stretchLinks.push_back(extk); BV[getNodeId(extk)] = 1;
          uint64_t len = 2;
          bool loop = (refk == extk);
          while (!loop && getNumActiveSuccessors(extk) == 1 && (!checkpredecessors || getNumActivePredecessors(extk) == 1)) {
            extk = getUniqueActiveSuccessor(extk);
            stretchLinks.push_back(extk);
            len += 1;
            int64_t extid = getNodeId(extk);
            if (BV[extid]) loop = true; else BV[extid] = 1;
          }
          uint64_t lastk = extk;
          for (uint64_t j = start; j < start + len; ++j) BV[getNodeId(stretchLinks[j])] = 0;
          if (loop && refk != lastk) {
            uint64_t j = 0;
            while (stretchLinks[start + j] != lastk) ++j;
            j += 1;
            uint64_t retract = len - j;
            len -= retract;
            stretchLinks.resize(stretchLinks.size() - retract);
          }
          Stretch s; s.first = refk; s.ext = firstext; s.last = lastk; s.len = len; s.stretchO = start;
          maxstretchlength = std::max(maxstretchlength, len);
          stretches.push_back(s);
        }
      }
    }
  }
  void stretchesUnique() {
    std::stable_sort(stretches.begin(), stretches.end());
    stretches.erase(std::unique(stretches.begin(), stretches.end()), stretches.end());
    uint64_t l = 0, o = 0;
    while (l < stretches.size()) {
      uint64_t h = l + 1;
      while (h < stretches.size() && stretches[h].first == stretches[l].first && stretches[h].ext == stretches[l].ext) ++h;
      stretches[o++] = stretches[l];
      l = h;
    }
    stretches.resize(o);
  }

  // ---- stretch position weights: :3176-3330 ----
  void computeFeasibleStretchPositions() {
    Astretchfeas.clear(); Acstretchfeas.clear();
    std::vector<std::vector<double>> buck;
    auto bpush = [&](uint64_t pp, double w) { if (pp >= buck.size()) buck.resize(pp + 1); buck[pp].push_back(w); };
    for (auto& st : stretches) {
      uint64_t len = st.len;
      for (int dir = 0; dir < 2; ++dir) {
        if (dir == 0) { st.feasposO = Astretchfeas.size(); st.feasposL = 0; } else { st.cfeasposO = Acstretchfeas.size(); st.cfeasposL = 0; }
        for (uint64_t jj = 0; jj < len; ++jj) {
          uint64_t j = dir == 0 ? jj : len - jj - 1;
          const Node* node = getNode(stretchLinks[st.stretchO + j]);
          uint64_t poff = len - jj - 1;
          if (dir == 0) for (uint64_t q = 0; q < node->numfeaspos; ++q) bpush(Afeaspos[node->feaspos + q].first + poff, Afeaspos[node->feaspos + q].second);
          else for (uint64_t q = 0; q < node->numcfeaspos; ++q) bpush(Acfeaspos[node->cfeaspos + q].first + poff, Acfeaspos[node->cfeaspos + q].second);
        }
        for (uint64_t zz = 0; zz < buck.size(); ++zz) {
          if (buck[zz].empty()) continue;
          if (buck[zz].size() == len && zz >= len - 1) {
            double weight = 0.0;
            for (uint64_t i = 0; i < len; ++i) weight += buck[zz][i];
            StretchFeasObject o{zz - (len - 1), weight, buck[zz][0], buck[zz][len - 1]};
            if (dir == 0) { Astretchfeas.push_back(o); st.feasposL += 1; } else { Acstretchfeas.push_back(o); st.cfeasposL += 1; }
          }
          buck[zz].clear();
        }
      }
    }
  }
  // :3388-3440
  double getReverseStretchLinkWeight(const Stretch& A, const Stretch& B) const {
    uint64_t shift = B.len - 1;
    double weight = 0.0;
    for (uint64_t ib = 0; ib < B.cfeasposL; ++ib) {
      const StretchFeasObject& ob = Acstretchfeas[B.cfeasposO + ib];
      for (uint64_t ia = 0; ia < A.cfeasposL; ++ia) {
        const StretchFeasObject& oa = Acstretchfeas[A.cfeasposO + ia];
        if (oa.p == ob.p + shift) { double lw = ob.w + (oa.w - oa.wf); weight = std::max(weight, lw); }
      }
    }
    return weight;
  }
  // :3442-3480
  void computeStretchLinks() {
    reverseStretchLinks.clear();
    for (uint64_t i = 0; i < stretches.size(); ++i) {
      auto er = std::equal_range(stretches.begin(), stretches.end(), stretches[i].last,
                                 FirstCmp());
      for (auto q = er.first; q != er.second; ++q) {
        double rweight = getReverseStretchLinkWeight(stretches[i], *q);
        if (rweight >= 1e-1) reverseStretchLinks.push_back({(uint64_t)(q - stretches.begin()), i});
      }
    }
    std::sort(reverseStretchLinks.begin(), reverseStretchLinks.end());
  }
  struct FirstCmp {
    bool operator()(const Stretch& a, uint64_t b) const { return a.first < b; }
    bool operator()(uint64_t a, const Stretch& b) const { return a < b.first; }
  };

  const StretchFeasObject* cachedStretchPositionWeight(uint64_t sid, uint64_t pp) const {           // :3906-3918
    const StretchFeasObject* a = Astretchfeas.data() + stretches[sid].feasposO; const StretchFeasObject* e = a + stretches[sid].feasposL;
    while (a != e && a->p < pp) ++a;
    return (a != e && a->p == pp) ? a : nullptr;
  }
  const StretchFeasObject* cachedStretchReversePositionWeight(uint64_t sid, uint64_t pp) const {    // :3920-3932
    const StretchFeasObject* a = Acstretchfeas.data() + stretches[sid].cfeasposO; const StretchFeasObject* e = a + stretches[sid].cfeasposL;
    while (a != e && a->p < pp) ++a;
    return (a != e && a->p == pp) ? a : nullptr;
  }
  Path extendPath(const Path P, uint64_t sid) {                 // :3934-3950, :3989-4056
    Path NP = P; NP.off = AP.size();
    for (uint64_t i = 0; i < P.len; ++i) { uint64_t v = AP[P.off + i]; AP.push_back(v); }
    const StretchFeasObject* SFO = cachedStretchPositionWeight(sid, P.pos);
    AP.push_back(sid);
    NP.len += 1;
    if (NP.len == 1) { NP.baselen = stretches[sid].len + kmersize - 1; NP.weight = SFO ? SFO->w : 0; }
    else { NP.baselen += stretches[sid].len - 1; if (SFO) NP.weight += SFO->w - SFO->wf; }
    NP.pos += stretches[sid].len - 1;
    return NP;
  }
  ReversePath extendReversePath(const ReversePath P, uint64_t sid) {   // :3952-3968, :4058-4105
    ReversePath NP = P; NP.linkoff = (uint32_t)APR.size();
    for (uint64_t i = 0; i < P.len; ++i) { uint64_t v = APR[P.linkoff + i]; APR.push_back(v); }
    const StretchFeasObject* SFO = cachedStretchReversePositionWeight(sid, P.pos);
    APR.push_back(sid);
    NP.len += 1;
    if (NP.len == 1) { NP.baselen = (uint16_t)(stretches[sid].len + kmersize - 1); NP.weight = SFO ? SFO->w : 0.0; }
    else { NP.baselen = (uint16_t)(NP.baselen + stretches[sid].len - 1); if (SFO) NP.weight += SFO->w - SFO->wf; }
    NP.pos = (uint16_t)(NP.pos + stretches[sid].len - 1);
    NP.front = (uint32_t)stretches[sid].first;
    return NP;
  }
  bool checkReversePathFeasiblePosition(const ReversePath& RP) const {   // :4130-4159
    if (!RP.len) return true;
    const Stretch& ls = stretches[APR[RP.linkoff + RP.len - 1]];
    uint64_t checkpos = RP.pos - (ls.len - 1);
    for (uint64_t i = 0; i < ls.cfeasposL; ++i) { const StretchFeasObject& S = Acstretchfeas[ls.cfeasposO + i]; if (S.p == checkpos && S.w >= 0.5) return true; }
    return false;
  }

  // :3541-3787
  void prepareTraverse(bool checkpredecessors, uint64_t first, uint64_t lastk, int64_t lmax) {
    computeStretches(checkpredecessors);
    splitStretches(first);
    splitStretches(lastk);
    stretchesUnique();
    computeFeasibleStretchPositions();
    computeStretchLinks();
    APR.clear(); ARP.clear();
    for (auto& h : ARPH) h.clear();
    if (getNode(lastk)) { ReversePath s; s.front = (uint32_t)lastk; s.baselen = (uint16_t)kmersize; RPST.push(s); }
    while (!RPST.empty()) {
      ReversePath RP = RPST.top(); RPST.popvoid();
      uint64_t srcbaselen = RP.baselen;
      while (!(srcbaselen < ARPH.size())) ARPH.emplace_back(REVERSE_PATH_HEAP_SIZE);
      if (ARPH[srcbaselen].full()) {
        if (RP.weight <= ARPH[srcbaselen].top().weight) continue;
        ARPH[srcbaselen].popvoid();
      }
      ARPH[srcbaselen].push(RP);
      ARP.push_back(RP);
      if (RP.len == 0) {
        for (uint64_t i = 0; i < stretches.size(); ++i)
          if (stretches[i].last == lastk) { ReversePath RPE = extendReversePath(RP, i); if (checkReversePathFeasiblePosition(RPE)) RPST.pushBump(RPE); }
      } else if (RP.baselen < (uint64_t)((lmax + 1) / 2)) {
        uint64_t laststretchid = APR[RP.linkoff + RP.len - 1];
        auto ep = std::equal_range(reverseStretchLinks.begin(), reverseStretchLinks.end(), std::pair<uint64_t, uint64_t>(laststretchid, 0),
                                   [](const std::pair<uint64_t, uint64_t>& a, const std::pair<uint64_t, uint64_t>& b) { return a.first < b.first; });
        for (auto q = ep.first; q != ep.second; ++q) {
          ReversePath RPE = extendReversePath(RP, q->second);
          if (checkReversePathFeasiblePosition(RPE)) RPST.pushBump(RPE);
        }
      }
    }
    // sort by (front, baselen), ties in insertion order (C7) : :3742
    std::stable_sort(ARP.begin(), ARP.end(), [](const ReversePath& a, const ReversePath& b) { return a.front != b.front ? a.front < b.front : a.baselen < b.baselen; });
    std::vector<std::pair<double, uint64_t>> ARWT(ARP.size());
    for (uint64_t i = 0; i < ARP.size(); ++i) ARWT[i] = {ARP[i].weight, i};
    std::sort(ARWT.begin(), ARWT.end());
    ARW.assign(ARP.size(), 0);
    for (uint64_t i = 0; i < ARP.size(); ++i) ARW[ARWT[i].second] = i;      // rank by (weight, index) ascending : :3744-3757
  }

  double getPairScore(const Path& P, const ReversePath& RP) const {          // :3482-3497
    uint64_t lsid = AP[P.off + P.len - 1];
    uint64_t lpos = P.pos - (stretches[lsid].len - 1);
    const StretchFeasObject* SFO = cachedStretchPositionWeight(lsid, lpos);
    return SFO ? (P.weight + RP.weight - SFO->wl) : (P.weight + RP.weight);
  }
  ScoreInterval getPrimaryScoreInterval(uint64_t left, uint64_t right, const Path& P) const {   // :3499-3511 (range max of ARW)
    uint64_t mi = left;
    for (uint64_t i = left; i < right; ++i) if (ARW[i] > ARW[mi]) mi = i;
    return ScoreInterval{left, right, mi, getPairScore(P, ARP[mi]), P};
  }
  bool nextScoreInterval(ScoreInterval& S) const {                           // :3513-3534 (range-previous-value on ARW)
    uint64_t v = ARW[S.current];
    if (!v) return false;
    int64_t best = -1; uint64_t bi = 0;
    for (uint64_t i = S.left; i < S.right; ++i) if (ARW[i] <= v - 1 && (int64_t)ARW[i] > best) { best = (int64_t)ARW[i]; bi = i; }
    if (best < 0) return false;
    S.current = bi; S.weight = getPairScore(S.P, ARP[S.current]);
    return true;
  }
  void consPushWord(uint64_t w) { for (unsigned i = 0, shift = 2 * (kmersize - 1); i < kmersize; ++i, shift -= 2) Acons.push_back(remapChar((w >> shift) & 3)); }
  void decodePathPair(const Path& P, const ReversePath& RP) {                // :4267-4300
    consPushWord(stretches[AP[P.off]].first);
    for (uint64_t i = 0; i < P.len; ++i) { const Stretch& s = stretches[AP[P.off + i]]; for (uint64_t j = 1; j < s.len; ++j) Acons.push_back(remapChar(stretchLinks[s.stretchO + j] & 3)); }
    for (uint64_t ii = 0; ii < RP.len; ++ii) {
      uint64_t i = RP.len - ii - 1;
      const Stretch& s = stretches[APR[RP.linkoff + i]];
      for (uint64_t j = 1; j < s.len; ++j) Acons.push_back(remapChar(stretchLinks[s.stretchO + j] & 3));
    }
  }

  void pushAPQ(const Path& P) {                                  // :4843-4862 == :4997-5017
    while (!(P.baselen < APQ.size())) APQ.emplace_back(PATH_HEAP_SIZE);
    auto& h = APQ[P.baselen];
    if (h.full()) { if (P.weight > h.top().weight) { h.popvoid(); h.push(P); } } else h.push(P);
  }

  // :4496-5170 (stretch based branch)
  bool traverse(int64_t lmin, int64_t lmax, const SeqRef* MA, uint64_t MAo, uint64_t /*maxfrontpath*/, uint64_t maxfullpath) {
    Acons.clear(); AP.clear(); ACC.clear(); CDH.clear();
    maxForPosList(0); maxLastList();
    uint64_t firstthres = maxFirst.size() ? (maxFirst[0].first * 3) / 4 : 0;
    uint64_t lastthres = maxLast.size() ? (maxLast[0].first * 3) / 4 : 0;
    const int64_t K = (int64_t)kmersize;
    for (uint64_t fi = 0; fi < maxFirst.size() && maxFirst[fi].first >= firstthres; ++fi)
      for (uint64_t li = 0; li < maxLast.size() && maxLast[li].first >= lastthres; ++li) {
        uint64_t first = maxFirst[fi].second, lastk = maxLast[li].second;
        prepareTraverse(true, first, lastk, lmax);
        AP.clear();
        for (uint64_t i = 0; i < stretches.size(); ++i) if (stretches[i].first == first) pushAPQ(extendPath(Path(), i));
        SIQ.clear();
        for (uint64_t zz = 0; zz < APQ.size(); ++zz)
          while (!APQ[zz].empty()) {
            const Path P = APQ[zz].top(); APQ[zz].popvoid();
            int64_t candlen = (int64_t)P.pos + K;
            uint64_t plast = stretches[AP[P.off + P.len - 1]].last;
            // reverse paths with front == plast and baselen in the window : :4919-4943
            auto lo = std::lower_bound(ARP.begin(), ARP.end(), plast, [](const ReversePath& a, uint64_t b) { return a.front < b; });
            auto hi = std::upper_bound(ARP.begin(), ARP.end(), plast, [](uint64_t a, const ReversePath& b) { return a < b.front; });
            uint16_t blo = (uint16_t)std::max(lmin + K - candlen, (int64_t)0);
            auto sub = std::lower_bound(lo, hi, blo, [](const ReversePath& a, uint16_t b) { return a.baselen < b; });
            uint16_t bhi = (uint16_t)std::max(lmax + K - candlen, (int64_t)0);
            auto sup = std::upper_bound(sub, hi, bhi, [](uint16_t a, const ReversePath& b) { return a < b.baselen; });
            if (sub != sup) SIQ.pushBump(getPrimaryScoreInterval(sub - ARP.begin(), sup - ARP.begin(), P));
            uint64_t plsid = AP[P.off + P.len - 1];
            auto er = std::equal_range(stretches.begin(), stretches.end(), stretches[plsid].last, FirstCmp());
            if (P.baselen < kmersize || ((int64_t)(P.baselen - kmersize) < ((lmax + 1) / 2))) {
              for (auto q = er.first; q != er.second; ++q) {
                uint64_t sid = q - stretches.begin();
                const StretchFeasObject* SFO = cachedStretchPositionWeight(sid, P.pos);
                double eweight = SFO ? SFO->w : 0.0;
                if (eweight > 0.1) {
                  Path EP = extendPath(P, sid);
                  if (EP.weight > 0.1 && (int64_t)(EP.pos + kmersize) <= lmax) pushAPQ(EP);
                }
              }
            }
          }
        uint64_t prevo = 0, prevlen = std::numeric_limits<uint64_t>::max();
        for (uint64_t nfp = 0; !SIQ.empty() && nfp < maxfullpath; ++nfp) {           // :5049-5092
          ScoreInterval SI = SIQ.top(); SIQ.popvoid();
          ScoreInterval SIC = SI;
          if (nextScoreInterval(SIC)) SIQ.pushBump(SIC);
          double weight = SI.weight;
          if (CDH.full()) { if (weight <= CDH.top().weight) continue; else CDH.popvoid(); }
          uint64_t consstart = Acons.size();
          decodePathPair(SI.P, ARP[SI.current]);
          uint64_t conslen = Acons.size() - consstart;
          if (conslen == prevlen && std::equal(Acons.begin() + prevo, Acons.begin() + prevo + prevlen, Acons.begin() + consstart)) continue;
          prevo = consstart; prevlen = conslen;
          ConsensusCandidate c; c.o = consstart; c.l = conslen; c.weight = weight; c.error = 0.0;
          CDH.push(c);
        }
      }
    CH.clear();                                                                     // :5101-5136
    while (!CDH.empty()) CH.pushBump(CDH.pop());
    while (!CH.empty()) {
      ConsensusCandidate CC = CH.pop();
      uint64_t e = 0;
      for (uint64_t j = 0; j < MAo; ++j) e += editDistance(Acons.data() + CC.o, CC.l, MA[j].first, MA[j].second);   // :5355-5363
      CC.error = (double)e;
      ACC.push_back(CC);
    }
    // :5156 (n <= 16 => libstdc++ insertion sort => stable; C7)
    std::stable_sort(ACC.begin(), ACC.end(), [](const ConsensusCandidate& a, const ConsensusCandidate& b) { return a.error < b.error; });
    return !ACC.empty();
  }
  // ---- error-profile estimator's trivial traversal: :1256-1278 (maxForPos), :1329-1357 (maxLastWord), :3794-3826 ----
  uint64_t maxForPos(uint64_t pos) const {
    uint64_t maxv = 0, maxc = 0;
    for (auto& node : nodes) {
      uint64_t c = 0;
      for (uint64_t j = 0; j < node.freq; ++j) if (SP[node.spo + j].pos == pos) ++c;
      if (c > maxc) { maxc = c; maxv = node.v; }
    }
    return maxv;
  }
  uint64_t maxLastWord() const {
    uint64_t maxv = 0, maxc = 0, l = 0;
    while (l < last.size()) {
      uint64_t h = l + 1;
      while (h < last.size() && (last[h] >> 32) == (last[l] >> 32)) ++h;
      if (h - l > maxc) { maxc = h - l; maxv = last[l] >> 32; }
      l = h;
    }
    return maxv;
  }
  // consensus = the one unitig leading from the most frequent first k-mer to the most frequent last k-mer
  bool traverseTrivial(std::string& cons) {
    cons.clear();
    const uint64_t first = maxForPos(0), lastk = maxLastWord();
    prepareTraverse(false, first, lastk, 0);
    for (auto& st : stretches)
      if (st.first == first && st.last == lastk) {
        for (unsigned i = 0, shift = 2 * (kmersize - 1); i < kmersize; ++i, shift -= 2) cons.push_back((char)remapChar((first >> shift) & 3));
        for (uint64_t j = 1; j < st.len; ++j) cons.push_back((char)remapChar(stretchLinks[st.stretchO + j] & 3));
        return true;
      }
    return false;
  }
  uint64_t getNumCandidates() const { return ACC.size(); }
  std::pair<const uint8_t*, const uint8_t*> getCandidate(uint64_t i) const { return {Acons.data() + ACC[i].o, Acons.data() + ACC[i].o + ACC[i].l}; }   // :5177-5182
  // :5476-5482, :5408-5447
  std::pair<uint64_t, uint64_t> checkCandidatesU(const SeqRef* I, uint64_t o) const {
    if (!getNumCandidates()) return {0, (uint64_t)std::numeric_limits<double>::max()};
    auto c = getCandidate(0); uint64_t e = 0;
    for (uint64_t j = 0; j < o; ++j) e += editDistance(c.first, c.second - c.first, I[j].first, I[j].second);
    return {0, e};
  }
  // src/DebruijnGraph.hpp:3794-3824 (used by the error-profile estimator only)
};

}  // namespace oracle
