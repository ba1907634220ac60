// daccord_main.cpp -- `daccord [options] reads.las reads.db` : drop-in command line for the window-consensus path
// of gt1/daccord (reference src/daccord.cpp:185-241 help text, :1061-1305 option handling, :2107-2540 main loop),
// with the per-window consensus running on a B200 through the C ABI in include/daccord_b200.h.
// Host side (this file + pile.hpp / vote.hpp): LAS + Dazzler-DB input, overlap selection, trace reconstruction,
// window/slice extraction, pile vote, FastA on stdout, progress on stderr.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <fstream>
#include <iostream>
#include <chrono>
#include <unistd.h>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "las.hpp"
#include "las_index.hpp"
#include "dazzdb.hpp"
#include "pile.hpp"
#include "vote.hpp"
#include "eprof.hpp"
#include "../../../include/daccord_b200.h"

using namespace dhost;

static const char* HELP =
    "usage: daccord [options] reads.las reads.db\n"
    "\t-t: number of host threads (default: all)\n\t-w: window size (default 40)\n\t-a: advance size (default 10)\n"
    "\t-d: max depth (default unlimited)\n\t-f: produce full sequences\n\t-V: verbosity\n\t-I: read interval i,j (inclusive)\n"
    "\t-J: reads part i,j\n\t-E: error profile file name (default input.las.eprof)\n\t--eprofonly: compute error profile only\n\t--keepeprof: keep an error profile older than the .las (default 1)\n\t-m: minimum window coverage (default 3)\n"
    "\t-e: maximum window error (default unlimited)\n\t-l: minimum length of output (default 0)\n"
    "\t--minfilterfreq: minimum k-mer filter frequency (default 0)\n\t--maxfilterfreq: maximum k-mer filter frequency (default 2)\n"
    "\t-D: maximum number of alignments considered per read (default 5000)\n\t-k: kmer size lo[,hi] (default 8)\n"
    "\t--device: CUDA device ordinal (default 0)\n\t--batchreads: A-reads per GPU batch (default 1024)\n\t--inflight: batches in flight on the GPU (default 3)\n";

struct Args { std::map<std::string, std::string> opt; std::vector<std::string> pos; };
static bool parse_args(int argc, char** argv, Args& A) {
  static const char* longs[] = {"minfilterfreq", "maxfilterfreq", "vard", "eprofonly", "deepprofileonly", "keepeprof", "device", "batchreads", "inflight", "help", "version", nullptr};
  int i = 1;
  for (; i < argc; ++i) {
    std::string s = argv[i];
    if (s.size() >= 2 && s[0] == '-' && s[1] == '-') {
      std::string body = s.substr(2), key, val; bool found = false;
      for (int j = 0; longs[j]; ++j) { size_t n = strlen(longs[j]); if (body.compare(0, n, longs[j]) == 0 && n > key.size()) { key = longs[j]; found = true; } }
      if (!found) { fprintf(stderr, "[E] unknown option %s\n", s.c_str()); return false; }
      val = body.substr(key.size());
      if (!val.empty() && val[0] == '=') val = val.substr(1);
      A.opt[key] = val;
    } else if (s.size() >= 2 && s[0] == '-' && !isdigit((unsigned char)s[1])) {
      A.opt[std::string(1, s[1])] = s.substr(2);          // libmaus2 ArgParser style: value attached (-w40)
    } else break;                                          // options must precede positionals (README.md:99)
  }
  for (; i < argc; ++i) A.pos.push_back(argv[i]);
  return true;
}
static bool parse_pair(const std::string& s, int64_t& a, int64_t& b) { return sscanf(s.c_str(), "%ld,%ld", &a, &b) == 2; }

int main(int argc, char** argv) {
  Args A;
  if (!parse_args(argc, argv, A)) return EXIT_FAILURE;
  if (A.opt.count("h") || A.opt.count("help") || A.pos.size() < 2) { fprintf(stderr, "%s", HELP); return A.pos.size() < 2 && !A.opt.count("h") && !A.opt.count("help") ? EXIT_FAILURE : EXIT_SUCCESS; }
  auto getu = [&](const char* k, uint64_t def) -> uint64_t { auto it = A.opt.find(k); return (it == A.opt.end() || it->second.empty()) ? def : strtoull(it->second.c_str(), nullptr, 10); };
  const std::string lasfn = A.pos[0], dbfn = A.pos[1];
  if (A.pos.size() > 2 && A.pos[2] != dbfn) { fprintf(stderr, "[E] asymmetric (DB1 != DB2) input is not supported by this build\n"); return EXIT_FAILURE; }
  if (A.opt.count("deepprofileonly")) { fprintf(stderr, "[E] --deepprofileonly is not supported by this build\n"); return EXIT_FAILURE; }
  if (getu("vard", 0)) { fprintf(stderr, "[E] --vard is not supported by this build\n"); return EXIT_FAILURE; }
  dcu_params prm; memset(&prm, 0, sizeof(prm));
  prm.w = (uint32_t)getu("w", 40); const uint32_t advance = (uint32_t)getu("a", 10);
  prm.k_lo = prm.k_hi = 8;
  if (A.opt.count("k")) { unsigned lo = 0, hi = 0; int n = sscanf(A.opt["k"].c_str(), "%u,%u", &lo, &hi); if (n < 1) { fprintf(stderr, "[E] unable to parse k argument %s\n", A.opt["k"].c_str()); return EXIT_FAILURE; } prm.k_lo = lo; prm.k_hi = n == 2 ? hi : lo; }
  prm.min_cov = (uint32_t)getu("m", 3); prm.max_err = getu("e", UINT64_MAX);
  prm.min_ff = (int32_t)getu("minfilterfreq", 0); prm.max_ff = (int32_t)getu("maxfilterfreq", 2);
  const uint64_t maxalign = getu("d", UINT64_MAX), maxinput = getu("D", 5000), minlen = getu("l", 0);
  const bool producefull = A.opt.count("f") != 0;
  int nthreads = (int)getu("t", 0);
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  const int device = (int)getu("device", 0); const uint64_t batchreads = getu("batchreads", 1024);
  auto t_start = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) { fprintf(stderr, "[T] %.3fs %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(), what); };
  try {
    PackedDB db; LasData las;
    fprintf(stderr, "[V] loading %s ...", dbfn.c_str()); read_dazzdb(dbfn, db); fprintf(stderr, "done.\n");
    // LAS: index of record offsets per A-read (cached as <las>.dcuidx; reference src/daccord.cpp:1075-1104), then only the shard's byte range
    LasIndex lidx; bool built = false;
    fprintf(stderr, "[V] indexing %s ...", lasfn.c_str()); get_las_index(lasfn, lidx, &built); fprintf(stderr, "%s.\n", built ? "done" : "cached");
    // read range: -J i,j or -I i,j (inclusive) (reference src/daccord.cpp:1116-1227)
    int64_t minaread = 0, maxaread = (int64_t)db.rlen.size() - 1;
    if (lidx.maxaread >= lidx.minaread) { minaread = lidx.minaread; maxaread = lidx.maxaread; } else maxaread = -1;
    if (A.opt.count("J")) {
      int64_t Icnt, Idiv; if (!parse_pair(A.opt["J"], Icnt, Idiv)) { fprintf(stderr, "[E] unable to parse %s\n", A.opt["J"].c_str()); return EXIT_FAILURE; }
      int64_t top = maxaread + 1, span = top > minaread ? top - minaread : 0;
      if (span && !Idiv) { fprintf(stderr, "[E] denominator of J argument cannot be zero\n"); return EXIT_FAILURE; }
      if (top > minaread) { int64_t part = Idiv ? (span + Idiv - 1) / Idiv : 0; int64_t lo = std::min(minaread + Icnt * part, top), hi = std::min(lo + part, top); if (hi > lo) { minaread = lo; maxaread = hi - 1; } else { minaread = 0; maxaread = -1; } }
    } else if (A.opt.count("I")) {
      int64_t lo, hi; if (!parse_pair(A.opt["I"], lo, hi)) { fprintf(stderr, "[E] unable to parse %s\n", A.opt["I"].c_str()); return EXIT_FAILURE; }
      minaread = std::max(lo, minaread); maxaread = std::min(hi, maxaread);
    }
    const int64_t toparead = maxaread >= 0 ? maxaread + 1 : maxaread;
    fprintf(stderr, "[V] minaread=%ld toparead=%ld\n", (long)minaread, (long)toparead);
    fprintf(stderr, "[V] loading %s ...", lasfn.c_str()); read_las_range(lasfn, lidx, minaread, toparead, las, nthreads); las.build_index(db.rlen.size()); validate_las(las, db.rlen);
    fprintf(stderr, "done (%zu of %ld overlaps).\n", las.ovl.size(), (long)lidx.novl);
    lap("inputs loaded");
    fprintf(stderr, "[V] minfilterfreq=%d maxfilterfreq=%d\n", prm.min_ff, prm.max_ff);
    // error profile: <las>.eprof or -E, in the reference's binary layout (eprof.hpp); estimated from the first <= 1024 A-reads when the
    // file does not exist, or is older than the .las and --keepeprof is off (reference src/daccord.cpp:1652-1880; keepeprof defaults to 1, :1303)
    std::string eproffn = A.opt.count("E") ? A.opt["E"] : lasfn + ".eprof";
    const bool keepeprof = getu("keepeprof", 1) != 0;
    uint64_t ecnt[4] = {0, 0, 0, 0};
    if (eprof_is_stale(eproffn, lasfn, keepeprof)) {
      ProfileCounts PC = estimate_profile(db, las, minaread, toparead, maxalign, maxinput, nthreads);
      fprintf(stderr, "usable=%lu unusable=%lu eavg=%g edif=%g\n", (unsigned long)PC.usable, (unsigned long)PC.unusable, PC.eavg, PC.edif);
      write_eprof(eproffn, PC.cnt, PC.eavg, PC.edif);
    }
    if (!read_eprof(eproffn, ecnt)) { fprintf(stderr, "[E] cannot read error profile %s\n", eproffn.c_str()); return EXIT_FAILURE; }
    const uint64_t em = ecnt[0], es = ecnt[1], ei = ecnt[2], ed = ecnt[3];
    if (em + es + ed == 0) { fprintf(stderr, "[E] error profile %s is empty (no usable window in the sampled reads)\n", eproffn.c_str()); return EXIT_FAILURE; }
    if (A.opt.count("eprofonly")) return EXIT_SUCCESS;                  // src/daccord.cpp:1895-1896
    const uint64_t len = em + es + ed, numerr = es + ed + ei;
    prm.p_i = (double)ei / (double)len; prm.p_d = (double)ed / (double)len; prm.est_cor = 1.0 - (double)numerr / (double)len;
    fprintf(stderr, "error estimates\nerate=%g\ncor=%g\nins=%g\ndel=%g\n", (double)numerr / len, prm.est_cor, prm.p_i, prm.p_d);
    fprintf(stderr, "[V] using kmer range [%u,%u]\n", prm.k_lo, prm.k_hi);
    // Main loop (replaces reference src/daccord.cpp:2107-2540).  The reference overlaps input, compute and ordered output over OpenMP
    // threads (:2107-2112, :2481-2534); here `inflight` worker threads each own a dcu_ctx and walk the batches round robin: while one
    // batch runs the window kernel, the others select overlaps, pile, vote and format (the library serialises only the window passes of
    // the contexts of one device).  Output is released in batch order, sequences numbered like -t1.
    const uint64_t maxalign_eff = std::min<uint64_t>(maxalign, 2047);       // deeper piles are capped like -d2047 (16-bit slice counts / 2048-entry KmerLimit table)
    if (maxalign != UINT64_MAX && maxalign > maxalign_eff) fprintf(stderr, "[W] -d%lu capped at %lu\n", (unsigned long)maxalign, (unsigned long)maxalign_eff);
    const int64_t nbatches = toparead > minaread ? (toparead - minaread + (int64_t)batchreads - 1) / (int64_t)batchreads : 0;
    int inflight = (int)getu("inflight", 3);
    if (inflight < 1) inflight = 1;
    if ((int64_t)inflight > nbatches) inflight = (int)std::max<int64_t>(1, nbatches);
    const int wthreads = std::max(1, nthreads / inflight);
    std::vector<dcu_ctx*> ctxs((size_t)inflight, nullptr);
    {   // one context per in-flight batch, created side by side (each builds its tables and allocates its buffers); the first one uploads the read
        // database, the others share it
      std::vector<int> rcs((size_t)inflight, 0);
      std::vector<std::thread> th;
      for (int i = 0; i < inflight; ++i) th.emplace_back([&, i]() {
        rcs[i] = dcu_create(&prm, device, &ctxs[i]);
        if (!rcs[i] && i == 0) rcs[i] = dcu_set_reads(ctxs[0], db.bytes.data(), db.bytes.size());
      });
      for (auto& t : th) t.join();
      for (int i = 0; i < inflight; ++i) {
        if (!rcs[i] && i > 0) rcs[i] = dcu_share_reads(ctxs[i], ctxs[0]);
        if (rcs[i]) { fprintf(stderr, "[E] dcu_create / dcu_set_reads: %s %s\n", dcu_strerror(rcs[i]), ctxs[i] ? dcu_last_error(ctxs[i]) : ""); return EXIT_FAILURE; }
      }
    }
    lap("contexts ready");
    PileParams PP; PP.w = prm.w; PP.a = advance; PP.maxalign = maxalign_eff; PP.maxinput = maxinput;
    VoteParams VP; VP.producefull = producefull; VP.minlen = minlen;
    uint64_t wellcounter = 0, totwin = 0, totok = 0, totlost = 0;
    std::mutex mu; std::condition_variable cv; int64_t turn = 0; std::atomic<int64_t> next{0}; std::atomic<bool> failed{false}; std::string failmsg;
    const bool gpu_pile = las.tspace <= 128 && !getenv("DACCORD_HOST_PILE");
    const bool gpu_vote = !getenv("DACCORD_HOST_VOTE");
    if (!gpu_pile && !getenv("DACCORD_HOST_PILE"))
      fprintf(stderr, "[W] trace reconstruction and slice extraction run on the host for this input (the GPU piler aligns trace tiles of at most 128 A bases; here tspace=%d)\n", las.tspace);
    auto fail = [&](const std::string& m) { std::lock_guard<std::mutex> g(mu); if (!failed.exchange(true)) failmsg = m; cv.notify_all(); };
    auto worker = [&](int wid) {
      dcu_ctx* ctx = ctxs[wid];
      std::vector<dcu_result> res; std::vector<uint8_t> cons, ops;
      for (;;) {
        const int64_t bi = next.fetch_add(1);
        if (bi >= nbatches || failed.load()) break;
        const int64_t b0 = minaread + bi * (int64_t)batchreads, b1 = std::min<int64_t>(b0 + (int64_t)batchreads, toparead), nr = b1 - b0;
        std::vector<dcu_window> win; std::vector<dcu_slice> sl; std::vector<uint64_t> first(nr + 1, 0);
        uint64_t nw = 0, ns = 0; int rc = 0;
        // stage 1: windows + slices -- on the GPU from the selected overlaps (dcu_pile), or by the host piler (dcu_upload)
        if (gpu_pile) {
          std::vector<std::vector<uint32_t>> sels(nr);
#pragma omp parallel for schedule(dynamic, 8) num_threads(wthreads)
          for (int64_t r = 0; r < nr; ++r) select_overlaps(las, (uint64_t)(b0 + r), maxinput, sels[r]);
          size_t tot = 0; for (auto& v : sels) tot += v.size();
          std::vector<dcu_overlap> ov; ov.reserve(tot);
          for (auto& sel : sels)
            for (auto i : sel) { const Overlap& o = las.ovl[i]; dcu_overlap x; memset(&x, 0, sizeof(x)); x.abpos = o.abpos; x.aepos = o.aepos; x.bbpos = o.bbpos; x.bepos = o.bepos; x.flags = o.flags; x.aread = o.aread; x.bread = o.bread; x.diffs = o.diffs; x.tlen = o.tlen; x.trace_off = o.trace_off; ov.push_back(x); }
          rc = dcu_pile(ctx, ov.data(), ov.size(), las.trace.data(), las.trace.size(), las.tspace, db.boff.data(), db.rlen.data(), db.rlen.size(), advance, maxalign_eff, &nw, &ns);
          if (rc) { fail(std::string("dcu_pile: ") + dcu_strerror(rc) + ": " + dcu_last_error(ctx)); break; }
        } else {
          std::vector<std::vector<dcu_window>> wv(nr); std::vector<std::vector<dcu_slice>> sv(nr);
#pragma omp parallel num_threads(wthreads)
          {
            ReadPiler RP(db, las, PP);
#pragma omp for schedule(dynamic, 1)
            for (int64_t i = 0; i < nr; ++i) {
              try { RP.pile((uint64_t)(b0 + i), wv[i], sv[i]); }
              catch (std::exception& e) {       // per-read failures are logged and skipped (reference src/daccord.cpp:2466-2478)
#pragma omp critical
                { fprintf(stderr, "[E] read %ld: %s\n", (long)(b0 + i), e.what()); }
                wv[i].clear(); sv[i].clear();
              }
            }
          }
          for (int64_t i = 0; i < nr; ++i) { uint32_t base = (uint32_t)sl.size(); for (auto x : wv[i]) { x.slice_begin += base; win.push_back(x); } sl.insert(sl.end(), sv[i].begin(), sv[i].end()); }
          nw = win.size(); ns = sl.size();
          rc = dcu_upload(ctx, win.data(), nw, sl.data(), ns);
          if (rc) { fail(std::string("dcu_upload: ") + dcu_strerror(rc) + ": " + dcu_last_error(ctx)); break; }
        }
        // stage 2: per-window consensus.  Windows beyond every capacity of the build come back as DCU_WIN_OVERFLOW: logged, treated as failed
        rc = dcu_launch(ctx, nullptr);
        if (rc) { fail(std::string("dcu_launch: ") + dcu_strerror(rc) + ": " + dcu_last_error(ctx)); break; }
        uint64_t second = 0, lost = 0; uint32_t d0 = 0, d1 = 0;
        dcu_last_stats2(ctx, &second, &lost, &d0, &d1);
        if (lost) fprintf(stderr, "[W] reads [%ld,%ld): %s (no consensus for these windows)\n", (long)b0, (long)b1, dcu_last_error(ctx));
        res.resize(nw);
        // stage 3: pile vote -- on the GPU (only corrected bases and the 16-byte results cross PCIe), or on the host from the full results
        std::vector<dcu_segment> seg; std::vector<char> chars; std::vector<std::string> parts;
        if (gpu_vote) {
          rc = dcu_download(ctx, res.data(), nullptr, nullptr);
          uint64_t nseg = 0, nch = 0;
          if (!rc) rc = dcu_vote(ctx, producefull ? 1 : 0, minlen, db.boff.data(), db.rlen.data(), db.rlen.size(), &nseg, &nch);
          seg.resize(nseg); chars.resize(nch + 1);
          if (!rc) rc = dcu_get_corrected(ctx, seg.data(), chars.data());
          if (rc) { fail(std::string("dcu_vote: ") + dcu_strerror(rc) + ": " + dcu_last_error(ctx)); break; }
        } else {
          cons.resize(nw * DCU_CONS_STRIDE); ops.resize(nw * DCU_OPS_STRIDE);
          rc = dcu_download(ctx, res.data(), cons.data(), ops.data());
          if (!rc && gpu_pile) { win.resize(nw); rc = dcu_get_windows(ctx, win.data(), nullptr); }
          if (rc) { fail(std::string("dcu_download: ") + dcu_strerror(rc) + ": " + dcu_last_error(ctx)); break; }
          { uint64_t wi = 0; for (int64_t i = 0; i < nr; ++i) { first[i] = wi; while (wi < nw && (int64_t)win[wi].aread == b0 + i) ++wi; } first[nr] = nw; }
          parts.resize(nr);
#pragma omp parallel for schedule(dynamic, 4) num_threads(wthreads)
          for (int64_t i = 0; i < nr; ++i) {
            if (first[i] == first[i + 1]) continue;
            std::vector<PileElement> PV;
            for (uint64_t wi = first[i]; wi < first[i + 1]; ++wi) if (res[wi].status == DCU_WIN_OK) place_window(win[wi], res[wi], cons.data() + wi * DCU_CONS_STRIDE, ops.data() + wi * DCU_OPS_STRIDE, PV);
            std::string ab;
            if (producefull) { std::vector<uint8_t> codes; decode_read(db, (uint32_t)(b0 + i), false, codes); ab.resize(codes.size()); for (size_t q = 0; q < codes.size(); ++q) ab[q] = "ACGT"[codes[q]]; }
            uint64_t c0 = 0; vote_read(b0 + i, PV, VP, ab, c0, parts[i]);
          }
        }
        uint64_t att = 0, okc = 0;
        for (auto& r : res) { att += r.status != DCU_WIN_SKIPPED; okc += r.status == DCU_WIN_OK; }
        // ordered release: batch bi writes after batch bi - 1
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return turn == bi || failed.load(); });
        if (failed.load()) break;
        if (gpu_vote) {
          std::string text; format_segments(seg.data(), seg.size(), chars.data(), wellcounter, text);
          fwrite(text.data(), 1, text.size(), stdout);
        } else {
          for (int64_t i = 0; i < nr; ++i) {        // sequences renumbered into the global counter
            const std::string& s = parts[i]; size_t p = 0;
            while (p < s.size()) {
              size_t e = s.find('\n', p); if (e == std::string::npos) e = s.size();
              if (s[p] == '>') { size_t s1 = s.find('/', p), s2 = s.find('/', s1 + 1); fwrite(s.data() + p, 1, s1 + 1 - p, stdout); fprintf(stdout, "%lu", (unsigned long)wellcounter++); fwrite(s.data() + s2, 1, e - s2, stdout); }
              else fwrite(s.data() + p, 1, e - p, stdout);
              fputc('\n', stdout); p = e + 1;
            }
          }
        }
        totwin += att; totok += okc; totlost += lost;
        fprintf(stderr, "[V] reads [%ld,%ld) windows %lu\n", (long)b0, (long)b1, (unsigned long)nw);
        ++turn;
        lk.unlock();
        cv.notify_all();
      }
    };
    {
      std::vector<std::thread> th;
      for (int i = 1; i < inflight; ++i) th.emplace_back(worker, i);
      worker(0);
      for (auto& t : th) t.join();
    }
    fflush(stdout);
    lap("batches done");
    if (failed.load()) { fprintf(stderr, "[E] %s\n", failmsg.c_str()); return EXIT_FAILURE; }
    if (totlost) fprintf(stderr, "[W] %lu windows exceeded the capacities of this build and have no consensus\n", (unsigned long)totlost);
    if (getenv("DACCORD_CLEAN_EXIT")) for (auto it = ctxs.rbegin(); it != ctxs.rend(); ++it) dcu_destroy(*it);      // the owner of the read database (ctxs[0]) last
    else {   // the output is complete: ending the process releases the device memory faster than freeing gigabytes of workspaces buffer by buffer
      double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
      fprintf(stderr, "[V] processed in time %.3fs, %lu windows attempted, %lu consensus\n", secs, (unsigned long)totwin, (unsigned long)totok);
      fflush(stderr);
      _exit(EXIT_SUCCESS);
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    fprintf(stderr, "[V] processed in time %.3fs, %lu windows attempted, %lu consensus\n", secs, (unsigned long)totwin, (unsigned long)totok);
  } catch (std::exception& e) { std::cerr << e.what() << std::endl; return EXIT_FAILURE; }
  return EXIT_SUCCESS;
}
