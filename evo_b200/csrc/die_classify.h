// Host-only part of the SM -> die calibration (die_map.cu): from the measured L2 latencies to a die label per SM.
// Kept free of CUDA so that tests/test_host.py can compile it with g++ and feed it synthetic latencies.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace evo {

// lat: sms x probes cycles per dependent L2 load; where: for CTA b of the pair probe, smid | cluster rank << 16.
// Fills die[0..sms) with 0 / 1 (SM 0 is on die 0) and returns nullptr, or returns why the measurement cannot be trusted.
inline const char* classify_dies(const std::vector<float>& lat, int sms, int PROBES, const std::vector<unsigned>& where, std::vector<int>& die) {
  const char* verdict = nullptr;
  bool ok = true;
  die.assign(sms, -1);
  // lat[s][p] = base + (per-SM offset) + (per-line offset) + gap/2 * die(s) * home(p) + noise, die and home = +-1.  Centre rows and
  // columns (removes both offsets and any clock drift between the SMs' turns) and take the leading singular vector by power
  // iteration: its sign per SM is the die.  Accepted only if that rank-1 term dominates and no SM sits near zero.
  if (ok) {
    std::vector<double> x((size_t)sms * PROBES), u(sms), v(PROBES);
    std::vector<double> row(sms, 0.0), col(PROBES, 0.0);
    double grand = 0.0;
    for (int s = 0; s < sms; ++s) for (int p = 0; p < PROBES; ++p) { const double t = lat[s * PROBES + p]; row[s] += t / PROBES; col[p] += t / sms; grand += t / ((double)sms * PROBES); }
    double total = 0.0;
    for (int s = 0; s < sms; ++s) for (int p = 0; p < PROBES; ++p) { const double t = lat[s * PROBES + p] - row[s] - col[p] + grand; x[(size_t)s * PROBES + p] = t; total += t * t; }
    for (int p = 0; p < PROBES; ++p) v[p] = x[p];                      // start from SM 0's row
    double sigma2 = 0.0;
    for (int it = 0; it < 40; ++it) {
      for (int s = 0; s < sms; ++s) { double a = 0.0; for (int p = 0; p < PROBES; ++p) a += x[(size_t)s * PROBES + p] * v[p]; u[s] = a; }
      double nu = 0.0; for (int s = 0; s < sms; ++s) nu += u[s] * u[s];
      nu = std::sqrt(nu); if (nu == 0.0) break;
      for (int s = 0; s < sms; ++s) u[s] /= nu;
      for (int p = 0; p < PROBES; ++p) { double a = 0.0; for (int s = 0; s < sms; ++s) a += x[(size_t)s * PROBES + p] * u[s]; v[p] = a; }
      double nv = 0.0; for (int p = 0; p < PROBES; ++p) nv += v[p] * v[p];
      sigma2 = nv; nv = std::sqrt(nv); if (nv == 0.0) break;
      for (int p = 0; p < PROBES; ++p) v[p] /= nv;
    }
    std::vector<double> mag(sms);
    for (int s = 0; s < sms; ++s) mag[s] = std::fabs(u[s]);
    std::nth_element(mag.begin(), mag.begin() + sms / 2, mag.end());
    const double med = mag[sms / 2];
    if (total <= 0.0 || sigma2 < 0.4 * total) { ok = false; verdict = "no dominant two-group structure in the L2 latencies"; }
    for (int s = 0; s < sms && ok; ++s) {
      if (std::fabs(u[s]) < 0.35 * med) { ok = false; verdict = "an SM belongs to neither group clearly"; }
      die[s] = (u[s] > 0) == (u[0] > 0) ? 0 : 1;
    }
  }
  int count[2] = {0, 0};
  if (ok) {
    for (int s = 0; s < sms; ++s) ++count[die[s]];
    for (int t = 0; t < sms / 2 && ok; ++t) if (die[2 * t] != die[2 * t + 1]) { ok = false; verdict = "the two SMs of a TPC are on different dies"; }
    if (ok && ((count[0] & 1) || count[0] < sms / 4 || count[1] < sms / 4)) { ok = false; verdict = "implausible split"; }
    // 2-CTA clusters: ranks 0 and 1 on SMs 2t and 2t+1 (either order) of one TPC, every TPC used once
    std::vector<int> used(sms / 2, 0);
    for (int b = 0; b < sms && ok; b += 2) {
      const unsigned a = where[b] & 0xffff, c = where[b + 1] & 0xffff;
      if (a >= (unsigned)sms || c >= (unsigned)sms || (a >> 1) != (c >> 1) || a == c || used[a >> 1]++) { ok = false; verdict = "a CTA pair does not sit on one TPC of its own"; }
    }
  }
  return ok ? nullptr : verdict;
}

}  // namespace evo
