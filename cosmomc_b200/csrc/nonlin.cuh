// K11 — non-linear lensing rescale of the sources and sigma_8, batched over parameter points: what CAMB
// does between the Boltzmann ODE output and the source spline whenever use_nonlinear_lensing = T (the batch3 default).
//
// Reference behaviour reproduced (paths relative to the reference root):
//   camb/modules.f90:1882-1948   Transfer_GetMatterPowerData: log P(k/h) = log(T^2 k pi 2pi h^3 P_s(k)) + natural spline in log k/h
//   camb/modules.f90:2033-2074   MatterPowerData_k (log-log cubic spline, linear extrapolation outside the table)
//   camb/modules.f90:2202-2268   Transfer_Get_SigmaR (R = 8 Mpc/h: sigma_8(z), running trapezoid in ln k)
//   camb/halofit_ppf.f90:96-352  NonLinear_GetNonLinRatios, wint, halofit (Takahashi 2012, the default), omega_m / omega_v
//   camb/cmbmain.f90:1145-1204   MakeNonlinearSources: Src(k, 3, tau) *= spline in tau of sqrt(P_NL / P_L)(k, z_i)
// Inputs from the CPU ODE stage: the matter transfer function T(k, z_i) at the NLL redshifts and their conformal times.
// Decomposition: CTA = (point, redshift) for the power table and the halofit search (the 3 000-point Gaussian-filter
// integral of every bisection step is spread over the CTA and reduced in a fixed order); thread = (point, wavenumber)
// for the rescale, walking the time samples of the resident source array (k fastest: coalesced).
#pragma once
#include "common.cuh"
#include "project.cuh"

namespace cb200 {

constexpr int NL_THREADS = 128;
constexpr int NL_MAXZ = 32;

struct NlParams {
  int np, n_kt, n_z;
  const double* initpower;  // [np][10]
  const double* cosmo;      // [np][6] h, omm0, omegav, fnu, w, wa
  const double* kh;         // [np][n_kt]
  const double* z;          // [n_z]
  const double* transfer;   // [np][n_z][n_kt]
  double* logkh;            // [np][n_kt]
  double* matpower;         // [np][n_z][n_kt]
  double* ddmat;            // [np][n_z][n_kt]
  double* ratio;            // [np][n_z][n_kt]
  double* spec;             // [np][n_z][3] rknl, rneff, rncur
  double* sigma8;           // [np][n_z]
  int* status;              // [np]
};

// MatterPowerData_k out of shared-memory copies of one redshift's table
__device__ __forceinline__ double nl_power_at(const double* lk, const double* m, const double* dd, int nk, double kh) {
  const double logk = log(kh);
  double out;
  if (logk < lk[0]) {
    const double dp = (m[1] - m[0]) / (lk[1] - lk[0]);
    out = m[0] + dp * (logk - lk[0]);
  } else if (logk > lk[nk - 1]) {
    const double dp = (m[nk - 1] - m[nk - 2]) / (lk[nk - 1] - lk[nk - 2]);
    out = m[nk - 1] + dp * (logk - lk[nk - 1]);
  } else {
    int lo = 0, hi = nk - 1;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (lk[mid] < logk) lo = mid; else hi = mid;
    }
    const double ho = lk[lo + 1] - lk[lo];
    const double a0 = (lk[lo + 1] - logk) / ho, b0 = 1 - a0;
    out = a0 * m[lo] + b0 * m[lo + 1] + ((a0 * a0 * a0 - a0) * dd[lo] + (b0 * b0 * b0 - b0) * dd[lo + 1]) * ho * ho / 6;
  }
  return exp(out);
}

__device__ __forceinline__ double nl_block_sum(double v, double* sh) {  // fixed-order sum over the CTA, result in every thread
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < NL_THREADS / 32; w++) s += sh[w];
  return s;
}

// sigma_8(z) of every point: thread = (point, redshift), the reference's running trapezoid over the table's wavenumbers
__global__ void nl_sigma8_kernel(NlParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.np * p.n_z) return;
  const int pt = t / p.n_z, itf = t - pt * p.n_z;
  const double* ip = p.initpower + (size_t)pt * 10;
  const double h = p.cosmo[(size_t)pt * 6];
  const double* kh = p.kh + (size_t)pt * p.n_kt;
  const double* tr = p.transfer + ((size_t)pt * p.n_z + itf) * p.n_kt;
  double sig8 = 0, dso = 0, lnko = 0;
  for (int ik = 0; ik < p.n_kt; ik++) {
    if (kh[ik] == 0) continue;
    const double k = kh[ik] * h;
    const double x = kh[ik] * 8.0;
    const double win = 3 * (sin(x) - x * cos(x)) / (x * x * x);
    const double lnk = log(k);
    const double dlnk = (ik == 0) ? 0.5 : lnk - lnko;
    const double d = (win * k * k) * (win * k * k) * scalar_power_dev(ip, k) * (tr[ik] * tr[ik]);
    sig8 = sig8 + (d + dso) * dlnk / 2;
    dso = d;
    lnko = lnk;
  }
  p.sigma8[t] = sqrt(sig8);
}

// power table + spline + halofit ratios: CTA = (point, redshift)
__global__ void __launch_bounds__(NL_THREADS) nl_halofit_kernel(NlParams p) {
  extern __shared__ __align__(16) double nl_smem[];
  __shared__ double sh[NL_THREADS / 32];
  const int nk = p.n_kt;
  double* lk = nl_smem;       // [nk]
  double* m = lk + nk;        // [nk]
  double* dd = m + nk;        // [nk]
  double* u = dd + nk;        // [nk] spline sweep
  const int pt = blockIdx.x / p.n_z, itf = blockIdx.x - pt * p.n_z, tid = threadIdx.x;
  const double* ip = p.initpower + (size_t)pt * 10;
  const double* cs = p.cosmo + (size_t)pt * 6;
  const double h = cs[0], omm0 = cs[1], omegav = cs[2], fnu = cs[3], w_hf = cs[4], wa_hf = cs[5];
  const double pi = kPi, twopi = kTwoPi;
  const double* kh = p.kh + (size_t)pt * nk;
  const double* tr = p.transfer + ((size_t)pt * p.n_z + itf) * nk;
  for (int ik = tid; ik < nk; ik += NL_THREADS) {
    const double k = kh[ik] * h;
    lk[ik] = log(kh[ik]);
    m[ik] = log(tr[ik] * tr[ik] * k * pi * twopi * (h * h * h) * scalar_power_dev(ip, k));
  }
  __syncthreads();
  if (tid == 0) {  // natural spline (camb/subroutines.f90:253-296, both end flags > 0.99e30)
    double d1r = (m[1] - m[0]) / (lk[1] - lk[0]), d1l;
    dd[0] = 0; u[0] = 0;
    for (int i = 1; i <= nk - 2; i++) {
      d1l = d1r;
      d1r = (m[i + 1] - m[i]) / (lk[i + 1] - lk[i]);
      const double xxdiv = 1 / (lk[i + 1] - lk[i - 1]);
      const double sig = (lk[i] - lk[i - 1]) * xxdiv;
      const double xp = 1 / (sig * dd[i - 1] + 2);
      dd[i] = (sig - 1) * xp;
      u[i] = (6 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
    }
    dd[nk - 1] = 0;
    for (int i = nk - 2; i >= 0; i--) dd[i] = dd[i] * dd[i + 1] + u[i];
  }
  __syncthreads();
  const size_t row = ((size_t)pt * p.n_z + itf) * nk;
  for (int ik = tid; ik < nk; ik += NL_THREADS) {
    p.matpower[row + ik] = m[ik]; p.ddmat[row + ik] = dd[ik];
    if (itf == 0) p.logkh[(size_t)pt * nk + ik] = lk[ik];
  }
  // ---- non-linear scale, effective index and curvature (Smith et al. 2002): bisection on the Gaussian-filter variance ----
  const double a = 1 / (1 + p.z[itf]);
  const double Qa2 = pow(a, -1.0 - 3.0 * (w_hf + wa_hf)) * exp(-3.0 * (1 - a) * wa_hf);
  const double omega_t = 1.0 + (omm0 + omegav - 1.0) / (1 - omm0 - omegav + omegav * Qa2 + omm0 / a);
  const double om_m = omega_t * omm0 / (omm0 + omegav * a * Qa2);
  const double om_v = omega_t * omegav * Qa2 / (omegav * Qa2 + omm0 / a);
  double xlogr1 = -2.0, xlogr2 = 3.5, rknl = 0, rneff = 0, rncur = 0;
  bool found = false, crazy = false;
  for (;;) {
    double rmid = (xlogr2 + xlogr1) / 2.0;
    rmid = pow(10.0, rmid);
    // wint: 3000 midpoints in t, k = 1/t - 1
    const int nint = 3000;
    const double anorm = 1 / (2 * pi * pi);
    double s1 = 0, s2 = 0, s3 = 0;
    for (int i = 1 + tid; i <= nint; i += NL_THREADS) {
      const double t = (i - 0.5) / nint;
      const double y = -1.0 + 1.0 / t;
      const double d2v = nl_power_at(lk, m, dd, nk, y) * (y * y * y * anorm);
      const double x = y * rmid, x2 = x * x;
      const double w1 = exp(-x2), w2 = 2 * x2 * w1, w3 = 4 * x2 * (1 - x2) * w1;
      const double fac = d2v / y / t / t;
      s1 += w1 * fac; s2 += w2 * fac; s3 += w3 * fac;
    }
    s1 = nl_block_sum(s1, sh) / nint; s2 = nl_block_sum(s2, sh) / nint; s3 = nl_block_sum(s3, sh) / nint;
    const double sig = sqrt(s1), d1 = -s2 / s1, d2 = -s2 * s2 / s1 / s1 - s3 / s1;
    const double diff = sig - 1.0;
    if (fabs(diff) <= 0.001) { rknl = 1. / rmid; rneff = -3 - d1; rncur = -d2; found = true; break; }
    else if (diff > 0.001) xlogr1 = log10(rmid);
    else if (diff < -0.001) xlogr2 = log10(rmid);
    if (xlogr2 < -1.9999) break;                         // still linear at this redshift
    else if (xlogr1 > 3.4999) { crazy = true; break; }   // "totally crazy non-linear": global_error_flag = 349
  }
  if (tid == 0) {
    double* sp = p.spec + ((size_t)pt * p.n_z + itf) * 3;
    sp[0] = rknl; sp[1] = rneff; sp[2] = rncur;
    if (crazy) p.status[pt] = 349;
  }
  // ---- halofit (Takahashi et al. 2012) ratio per wavenumber ----
  const double rn = rneff;
  const double gam = 0.1971 - 0.0843 * rn + 0.8460 * rncur;
  const double de = om_v * (1. + w_hf + wa_hf * (1 - a));
  const double ca = pow(10.0, 1.5222 + 2.8553 * rn + 2.3706 * rn * rn + 0.9903 * rn * rn * rn + 0.2250 * rn * rn * rn * rn - 0.6038 * rncur + 0.1749 * de);
  const double cb = pow(10.0, -0.5642 + 0.5864 * rn + 0.5716 * rn * rn - 1.5474 * rncur + 0.2279 * de);
  const double cc = pow(10.0, 0.3698 + 2.0404 * rn + 0.8161 * rn * rn + 0.5869 * rncur);
  const double xnu = pow(10.0, 5.2105 + 3.6902 * rn);
  const double alpha = fabs(6.0835 + 1.3373 * rn - 0.1959 * rn * rn - 5.5274 * rncur);
  const double beta = 2.0379 - 0.7354 * rn + 0.3157 * rn * rn + 1.2490 * rn * rn * rn + 0.3980 * rn * rn * rn * rn - 0.1682 * rncur +
                      fnu * (1.081 + 0.395 * rn * rn);
  double f1 = 1, f2 = 1, f3 = 1;
  if (fabs(1 - om_m) > 0.01) {
    const double frac = om_v / (1. - om_m);
    f1 = frac * pow(om_m, -0.0307) + (1 - frac) * pow(om_m, -0.0732);
    f2 = frac * pow(om_m, -0.0585) + (1 - frac) * pow(om_m, -0.1423);
    f3 = frac * pow(om_m, 0.0743) + (1 - frac) * pow(om_m, 0.0725);
  }
  for (int ik = tid; ik < nk; ik += NL_THREADS) {
    double r = 1.0;
    const double rk = exp(lk[ik]);
    if (found && rk > (double)0.005f) {
      const double plin = nl_power_at(lk, m, dd, nk, rk) * (rk * rk * rk / (2 * pi * pi));
      const double y = rk / rknl;
      double ph = ca * pow(y, f1 * 3) / (1 + cb * pow(y, f2) + pow(f3 * cc * y, 3 - gam));
      ph = ph / (1 + xnu / (y * y)) * (1 + fnu * 0.977);
      const double plinaa = plin * (1 + fnu * 47.48 * rk * rk / (1 + 1.5 * rk * rk));
      const double pq = plin * pow(1 + plinaa, beta) / (1 + plinaa * alpha) * exp(-y / 4.0 - y * y / 8.0);
      r = sqrt((pq + ph) / plin);
    }
    p.ratio[row + ik] = r;
  }
}

// MakeNonlinearSources: thread = (point, source wavenumber); the first n_k transfer wavenumbers are the source wavenumbers
__global__ void nl_rescale_kernel(PointView v, int p0, int np, int n_kt, int n_z, const double* __restrict__ cosmo,
                                  const double* __restrict__ ratio, const double* __restrict__ tautf, double* __restrict__ src) {
  const int lp = blockIdx.y, ik = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np) return;
  const int pt = p0 + lp;
  const int nk = v.n_k[pt], nt = v.n_tau[pt];
  if (ik >= nk) return;
  const double h = cosmo[(size_t)lp * 6];
  const double kq = v.ksrc[(size_t)pt * v.NK + ik];
  if (!(kq / h > (double)0.005f)) return;
  double sc[NL_MAXZ], d2[NL_MAXZ], u[NL_MAXZ];
  const double* tf = tautf + (size_t)lp * n_z;
  bool all_small = true;
  for (int i = 0; i < n_z; i++) {
    sc[i] = ratio[((size_t)lp * n_z + i) * n_kt + ik];
    if (!(fabs(sc[i] - 1) < 5e-4)) all_small = false;
  }
  if (all_small) return;
  {  // natural spline of the scaling over the transfer times (spl_large end flags)
    double d1r = (sc[1] - sc[0]) / (tf[1] - tf[0]), d1l;
    d2[0] = 0; u[0] = 0;
    for (int i = 1; i <= n_z - 2; i++) {
      d1l = d1r;
      d1r = (sc[i + 1] - sc[i]) / (tf[i + 1] - tf[i]);
      const double xxdiv = 1 / (tf[i + 1] - tf[i - 1]);
      const double sig = (tf[i] - tf[i - 1]) * xxdiv;
      const double xp = 1 / (sig * d2[i - 1] + 2);
      d2[i] = (sig - 1) * xp;
      u[i] = (6 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
    }
    d2[n_z - 1] = 0;
    for (int i = n_z - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
  }
  const double* tau = v.tau + (size_t)pt * v.NT;
  double* s = src + ((size_t)pt * v.NT * v.NSRC + 2) * v.NK + ik;   // source 3 (lensing potential) of time sample 1
  int tf_lo = 0;
  for (int i = 1; i <= nt - 1; i++) {
    const double t = tau[i - 1];
    if (t < tf[0]) continue;
    while (t > tf[tf_lo + 1]) tf_lo++;
    const double ho = tf[tf_lo + 1] - tf[tf_lo];
    const double a0 = (tf[tf_lo + 1] - t) / ho, b0 = 1 - a0;
    const double ascale = a0 * sc[tf_lo] + b0 * sc[tf_lo + 1] + ((a0 * a0 * a0 - a0) * d2[tf_lo] + (b0 * b0 * b0 - b0) * d2[tf_lo + 1]) * ho * ho / 6;
    s[(size_t)(i - 1) * v.NSRC * v.NK] *= ascale;
  }
}

}  // namespace cb200
