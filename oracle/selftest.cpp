// TEST INFRASTRUCTURE: quick sanity run of the oracle on random windows (accuracy vs known truth)
#include "window.hpp"
#include <random>
#include <cstdio>
#include <chrono>
using namespace oracle;
static std::string noisy(const std::string& t, std::mt19937_64& g, double pi, double pd, double ps) {
  std::uniform_real_distribution<double> U(0, 1); std::string r;
  for (size_t i = 0; i < t.size(); ++i) {
    while (U(g) < pi) r.push_back("ACGT"[g() & 3]);
    double x = U(g);
    if (x < pd) continue;
    if (x < pd + ps) { char c; do c = "ACGT"[g() & 3]; while (c == t[i]); r.push_back(c); } else r.push_back(t[i]);
  }
  return r;
}
int main(int argc, char** argv) {
  int nwin = argc > 1 ? atoi(argv[1]) : 200; int depth = argc > 2 ? atoi(argv[2]) : 40;
  Params P; Tables T(P); WindowContext C(T);
  std::mt19937_64 g(12345);
  uint64_t ok = 0, exact = 0, att = 0, ed = 0, tl = 0; uint64_t ffh[4] = {0, 0, 0, 0};
  auto t0 = std::chrono::steady_clock::now();
  for (int w = 0; w < nwin; ++w) {
    std::string truth; for (int i = 0; i < 40; ++i) truth.push_back("ACGT"[g() & 3]);
    std::vector<std::string> S; std::string a = noisy(truth, g, 0.09, 0.045, 0.015);
    a.resize(40, 'A'); S.push_back(a);
    for (int j = 0; j < depth; ++j) S.push_back(noisy(truth, g, 0.09, 0.045, 0.015));
    std::vector<SeqRef> MA; for (auto& s : S) MA.push_back({(const uint8_t*)s.data(), s.size()});
    WindowResult R = C.run(MA.data(), MA.size());
    att += R.attempted; ok += R.ok;
    if (R.ok) { exact += (R.cons == truth); ed += editDistance((const uint8_t*)R.cons.data(), R.cons.size(), (const uint8_t*)truth.data(), truth.size()); tl += truth.size(); ffh[R.filterfreq + 1]++; }
    if (w < 3) printf("truth %s\ncons  %s elen=%ld ff=%ld err=%lu ncand=%lu\n", truth.c_str(), R.cons.c_str(), (long)R.elength, (long)R.filterfreq, (unsigned long)R.minrate, (unsigned long)R.ncand);
  }
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("windows %d attempted %lu ok %lu exact %lu cons-vs-truth erate %.5f  ff2/1/0 = %lu/%lu/%lu  %.1f win/s\n", nwin, (unsigned long)att, (unsigned long)ok, (unsigned long)exact, tl ? (double)ed / tl : 0.0, (unsigned long)ffh[3], (unsigned long)ffh[2], (unsigned long)ffh[1], nwin / dt);
  return 0;
}
