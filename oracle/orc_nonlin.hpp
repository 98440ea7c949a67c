// ORACLE (test infrastructure, NOT product code): CPU restatement of CAMB's non-linear lensing rescale and sigma_8
// (SURVEY 8f-2), the step between the Boltzmann ODE output and the source spline whenever use_nonlinear_lensing = T.
//
//   Transfer_GetMatterPowerData, MatterPowerdata_getsplines, MatterPowerData_k   camb/modules.f90:1882-2074
//   Transfer_Get_SigmaR (sigma_8 = R 8 Mpc/h, trapezoid in ln k)                 camb/modules.f90:2202-2268
//   NonLinear_GetNonLinRatios, halofit (Takahashi 2012), wint, omega_m/omega_v    camb/halofit_ppf.f90:96-352
//   MakeNonlinearSources                                                         camb/cmbmain.f90:1145-1204
//
// Parity unpinned by reference output (no matter transfer functions can be produced here: they come from the ODE stage
// of the Fortran build); pinned instead to known answers that do not depend on this restatement
// (tests/test_nonlinear_oracle.py): a pure power law has n_eff = n, curvature 0 and an analytic non-linear scale, and
// sigma_R converges to the quadrature of the same integrand.
#pragma once
#include <cmath>
#include <vector>

#include "orc_core.hpp"

namespace orc {

struct MatterPower {  // MatterPowerData (camb/modules.f90:1800-1815) for one point
  int num_k = 0, num_z = 0;
  std::vector<double> log_kh;                 // [num_k]
  std::vector<std::vector<double>> matpower;  // [num_z][num_k] log P(k/h) in (Mpc/h)^3
  std::vector<std::vector<double>> ddmat;
  std::vector<std::vector<double>> nonlin_ratio;
  std::vector<double> redshifts;

  // Transfer_GetMatterPowerData: P = T^2 k pi 2pi h^3 P_s(k); transfer [num_z][num_k], kh [num_k]
  template <class PS>
  void from_transfer(int nk, int nz, const double* kh, const double* transfer, const double* z, double h, PS scalar_power) {
    const double pi = 3.1415926535897932384626433832795, twopi = 2 * pi;
    num_k = nk; num_z = nz;
    log_kh.resize(nk); redshifts.assign(z, z + nz);
    matpower.assign(nz, std::vector<double>(nk)); ddmat.assign(nz, std::vector<double>(nk));
    nonlin_ratio.assign(nz, std::vector<double>(nk, 1.0));
    for (int ik = 0; ik < nk; ik++) {
      const double k = kh[ik] * h;
      log_kh[ik] = std::log(kh[ik]);
      const double power = scalar_power(k);
      for (int itf = 0; itf < nz; itf++) {
        const double t = transfer[(size_t)itf * nk + ik];
        matpower[itf][ik] = std::log(t * t * k * pi * twopi * (h * h * h) * power);
      }
    }
    for (int itf = 0; itf < nz; itf++) spline(log_kh.data(), matpower[itf].data(), nk, 1e30, 1e30, ddmat[itf].data());
  }
  // MatterPowerData_k (camb/modules.f90:2033-2074): log-log cubic spline, linear extrapolation outside the table
  double at(double kh, int itf) const {
    const double logk = std::log(kh);
    const std::vector<double>& m = matpower[itf];
    double out;
    if (logk < log_kh[0]) {
      const double dp = (m[1] - m[0]) / (log_kh[1] - log_kh[0]);
      out = m[0] + dp * (logk - log_kh[0]);
    } else if (logk > log_kh[num_k - 1]) {
      const double dp = (m[num_k - 1] - m[num_k - 2]) / (log_kh[num_k - 1] - log_kh[num_k - 2]);
      out = m[num_k - 1] + dp * (logk - log_kh[num_k - 1]);
    } else {
      int llo = 0;  // last index with log_kh[llo] <= logk (the reference's running search gives the same bracket)
      int lo = 0, hi = num_k - 1;
      while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (log_kh[mid] < logk) lo = mid; else hi = mid; }
      llo = lo;
      const int lhi = llo + 1;
      const double ho = log_kh[lhi] - log_kh[llo];
      const double a0 = (log_kh[lhi] - logk) / ho, b0 = 1 - a0;
      out = a0 * m[llo] + b0 * m[lhi] + ((a0 * a0 * a0 - a0) * ddmat[itf][llo] + (b0 * b0 * b0 - b0) * ddmat[itf][lhi]) * ho * ho / 6;
    }
    return std::exp(out);
  }
};

// Transfer_Get_SigmaR (camb/modules.f90:2202-2268): sigma_R(z) for every redshift, R in Mpc/h
template <class PS>
inline void sigma_R(int nk, int nz, const double* kh, const double* transfer, double h, double R, PS scalar_power, double* out) {
  std::vector<double> sig8(nz, 0.0), dsig8o(nz, 0.0);
  double lnko = 0;
  for (int ik = 0; ik < nk; ik++) {
    if (kh[ik] == 0) continue;
    const double k = kh[ik] * h;
    const double x = kh[ik] * R;
    const double win = 3 * (std::sin(x) - x * std::cos(x)) / (x * x * x);
    const double lnk = std::log(k);
    const double dlnk = (ik == 0) ? 0.5 : lnk - lnko;
    const double powers = scalar_power(k);
    for (int itf = 0; itf < nz; itf++) {
      const double t = transfer[(size_t)itf * nk + ik];
      const double d = (win * k * k) * (win * k * k) * powers * (t * t);
      sig8[itf] = sig8[itf] + (d + dsig8o[itf]) * dlnk / 2;
      dsig8o[itf] = d;
    }
    lnko = lnk;
  }
  for (int itf = 0; itf < nz; itf++) out[itf] = std::sqrt(sig8[itf]);
}

struct Halofit {  // camb/halofit_ppf.f90:96-352, halofit_version = takahashi (the default)
  double omm0 = 0, omegav = 0, fnu = 0, w_lam = -1, wa = 0;
  double om_m = 0, om_v = 0, acur = 1;
  static double omega_m(double aa, double om_m0, double om_v0, double w, double wa_) {
    const double Qa2 = std::pow(aa, -1.0 - 3.0 * (w + wa_)) * std::exp(-3.0 * (1 - aa) * wa_);
    const double omega_t = 1.0 + (om_m0 + om_v0 - 1.0) / (1 - om_m0 - om_v0 + om_v0 * Qa2 + om_m0 / aa);
    return omega_t * om_m0 / (om_m0 + om_v0 * aa * Qa2);
  }
  static double omega_v(double aa, double om_m0, double om_v0, double w, double wa_) {
    const double Qa2 = std::pow(aa, -1.0 - 3.0 * (w + wa_)) * std::exp(-3.0 * (1 - aa) * wa_);
    const double omega_t = 1.0 + (om_m0 + om_v0 - 1.0) / (1 - om_m0 - om_v0 + om_v0 * Qa2 + om_m0 / aa);
    return omega_t * om_v0 * Qa2 / (om_v0 * Qa2 + om_m0 / aa);
  }
  static void wint(const MatterPower& PK, int itf, double r, double& sig, double& d1, double& d2) {
    const double pi = 3.1415926535897932384626433832795;
    const int nint = 3000;
    double sum1 = 0, sum2 = 0, sum3 = 0;
    const double anorm = 1 / (2 * pi * pi);
    for (int i = 1; i <= nint; i++) {
      const double t = (i - 0.5) / nint;
      const double y = -1.0 + 1.0 / t;
      const double rk = y;
      const double dd = PK.at(rk, itf) * (rk * rk * rk * anorm);
      const double x = y * r, x2 = x * x;
      const double w1 = std::exp(-x2), w2 = 2 * x2 * w1, w3 = 4 * x2 * (1 - x2) * w1;
      const double fac = dd / y / t / t;
      sum1 = sum1 + w1 * fac; sum2 = sum2 + w2 * fac; sum3 = sum3 + w3 * fac;
    }
    sum1 = sum1 / nint; sum2 = sum2 / nint; sum3 = sum3 / nint;
    sig = std::sqrt(sum1);
    d1 = -sum2 / sum1;
    d2 = -sum2 * sum2 / sum1 / sum1 - sum3 / sum1;
  }
  void halofit(double rk, double rn, double rncur, double rknl, double plin, double& pnl) const {
    const double gam = 0.1971 - 0.0843 * rn + 0.8460 * rncur;
    double a = 1.5222 + 2.8553 * rn + 2.3706 * rn * rn + 0.9903 * rn * rn * rn + 0.2250 * rn * rn * rn * rn - 0.6038 * rncur +
               0.1749 * om_v * (1. + w_lam + wa * (1 - acur));
    a = std::pow(10.0, a);
    const double b = std::pow(10.0, -0.5642 + 0.5864 * rn + 0.5716 * rn * rn - 1.5474 * rncur + 0.2279 * om_v * (1. + w_lam + wa * (1 - acur)));
    const double c = std::pow(10.0, 0.3698 + 2.0404 * rn + 0.8161 * rn * rn + 0.5869 * rncur);
    const double xmu = 0., xnu = std::pow(10.0, 5.2105 + 3.6902 * rn);
    const double alpha = std::fabs(6.0835 + 1.3373 * rn - 0.1959 * rn * rn - 5.5274 * rncur);
    const double beta = 2.0379 - 0.7354 * rn + 0.3157 * rn * rn + 1.2490 * rn * rn * rn + 0.3980 * rn * rn * rn * rn - 0.1682 * rncur +
                        fnu * (1.081 + 0.395 * rn * rn);
    double f1, f2, f3;
    if (std::fabs(1 - om_m) > 0.01) {
      const double f1a = std::pow(om_m, -0.0732), f2a = std::pow(om_m, -0.1423), f3a = std::pow(om_m, 0.0725);
      const double f1b = std::pow(om_m, -0.0307), f2b = std::pow(om_m, -0.0585), f3b = std::pow(om_m, 0.0743);
      const double frac = om_v / (1. - om_m);
      f1 = frac * f1b + (1 - frac) * f1a; f2 = frac * f2b + (1 - frac) * f2a; f3 = frac * f3b + (1 - frac) * f3a;
    } else { f1 = 1; f2 = 1; f3 = 1; }
    const double y = rk / rknl;
    double ph = a * std::pow(y, f1 * 3) / (1 + b * std::pow(y, f2) + std::pow(f3 * c * y, 3 - gam));
    ph = ph / (1 + xmu / y + xnu / (y * y)) * (1 + fnu * 0.977);
    const double plinaa = plin * (1 + fnu * 47.48 * rk * rk / (1 + 1.5 * rk * rk));
    const double pq = plin * std::pow(1 + plinaa, beta) / (1 + plinaa * alpha) * std::exp(-y / 4.0 - y * y / 8.0);
    pnl = pq + ph;
  }
  // NonLinear_GetNonLinRatios: fills PK.nonlin_ratio; spec[nz][3] = rknl, rneff, rncur (0 where still linear); returns 349
  // for the "totally crazy non-linear" exit
  int ratios(MatterPower& PK, double* spec) {
    const double pi = 3.1415926535897932384626433832795;
    const double Min_kh_nonlinear = (double)0.005f;
    int err = 0;
    for (int itf = 0; itf < PK.num_z; itf++) {
      for (double& v : PK.nonlin_ratio[itf]) v = 1;
      if (spec) spec[itf * 3] = spec[itf * 3 + 1] = spec[itf * 3 + 2] = 0;
      const double a = 1 / (1 + PK.redshifts[itf]);
      om_m = omega_m(a, omm0, omegav, w_lam, wa);
      om_v = omega_v(a, omm0, omegav, w_lam, wa);
      acur = a;
      double xlogr1 = -2.0, xlogr2 = 3.5, rknl = 0, rneff = 0, rncur = 0;
      bool found = false;
      for (;;) {
        double rmid = (xlogr2 + xlogr1) / 2.0;
        rmid = std::pow(10.0, rmid);
        double sig, d1, d2;
        wint(PK, itf, rmid, sig, d1, d2);
        const double diff = sig - 1.0;
        if (std::fabs(diff) <= 0.001) { rknl = 1. / rmid; rneff = -3 - d1; rncur = -d2; found = true; break; }
        else if (diff > 0.001) xlogr1 = std::log10(rmid);
        else if (diff < -0.001) xlogr2 = std::log10(rmid);
        if (xlogr2 < -1.9999) break;                       // still linear
        else if (xlogr1 > 3.4999) { err = 349; break; }    // totally crazy non-linear
      }
      if (!found) continue;
      if (spec) { spec[itf * 3] = rknl; spec[itf * 3 + 1] = rneff; spec[itf * 3 + 2] = rncur; }
      for (int i = 0; i < PK.num_k; i++) {
        const double rk = std::exp(PK.log_kh[i]);
        if (rk > Min_kh_nonlinear) {
          const double plin = PK.at(rk, itf) * (rk * rk * rk / (2 * pi * pi));
          double pnl;
          halofit(rk, rneff, rncur, rknl, plin, pnl);
          PK.nonlin_ratio[itf][i] = std::sqrt(pnl / plin);
        }
      }
    }
    return err;
  }
};

// MakeNonlinearSources (camb/cmbmain.f90:1145-1204): scales Src(ik, 3, i) (k fastest: src[tau][3][n_k]) in place.
// tau [n_tau] = TimeSteps%points, tautf [nz] ascending conformal times of the transfer redshifts, k [n_k] = Evolve_q.
inline void make_nonlinear_sources(const MatterPower& PK, int n_k, const double* k, double h, int n_tau, const double* tau,
                                   const double* tautf, double* src) {
  const double Min_kh_nonlinear = (double)0.005f;
  const int nz = PK.num_z;
  int first_step = 1;
  while (tau[first_step - 1] < tautf[0]) first_step++;
  std::vector<double> scaling(nz), dd(nz);
  for (int ik = 0; ik < n_k; ik++) {
    if (!(k[ik] / h > Min_kh_nonlinear)) continue;
    bool all_small = true;
    for (int i = 0; i < nz; i++) { scaling[i] = PK.nonlin_ratio[i][ik]; if (!(std::fabs(scaling[i] - 1) < 5e-4)) all_small = false; }
    if (all_small) continue;
    spline(tautf, scaling.data(), nz, 1e40, 1e40, dd.data());
    int tf_lo = 1, tf_hi = 2;
    for (int i = first_step; i <= n_tau - 1; i++) {
      const double t = tau[i - 1];
      while (t > tautf[tf_hi - 1]) { tf_lo++; tf_hi++; }
      const double ho = tautf[tf_hi - 1] - tautf[tf_lo - 1];
      const double a0 = (tautf[tf_hi - 1] - t) / ho, b0 = 1 - a0;
      const double ascale = a0 * scaling[tf_lo - 1] + b0 * scaling[tf_hi - 1] +
                            ((a0 * a0 * a0 - a0) * dd[tf_lo - 1] + (b0 * b0 * b0 - b0) * dd[tf_hi - 1]) * ho * ho / 6;
      src[((size_t)(i - 1) * 3 + 2) * n_k + ik] *= ascale;
    }
  }
}

}  // namespace orc
