// K5 — background functions (distances, H(z), theta, age) batched over parameter points, and the
// background-only likelihoods that consume them (BAO, MGS, HST, supernova set-up/finish).
//
// Reference behaviour reproduced (paths relative to the reference root):
//   camb/modules.f90:335-375      CAMBParams_Set densities (grhom, grhog, grhor, ...)
//   camb/modules.f90:1532-1718    MassiveNu: Nu_init / nuRhoPres / Nu_rho (log-spaced spline table of rho_nu(a m))
//   camb/subroutines.f90:6-50,342-364  splini / splder / splint
//   camb/equations.f90:69-100     dtauda
//   camb/subroutines.f90:117-176  rombint (same trapezoid/Richardson sequence, same stopping rule)
//   camb/modules.f90:519-751      DeltaTime, ComovingRadialDistance, AngularDiameterDistance, Hofz, BAO_D_v,
//                                 dsound_da, CosmomcTheta, DeltaPhysicalTimeGyr
//   source/bao.f90:265-308,390-410  BAO_LnLike, BAO_MGS_loglike ; source/HST.f90:47-59 HST_LnLike
// One thread per (parameter point, redshift); lanes of a warp share the redshift so that the Romberg loops of a
// warp run the same number of refinements.
#pragma once
#include "common.cuh"

namespace cb200 {

constexpr int NBG = 16;  // bg vector length (cosmomc_b200/params.py)
constexpr int NU_NRHOPN = 2000;

namespace bgc {  // camb/constants.f90
constexpr double pi = 3.1415926535897932384626433832795;
constexpr double c = 2.99792458e8;
constexpr double G = 6.6738e-11;
constexpr double sigma_boltz = 5.6704e-8;
constexpr double Gyr = 3.1556926e16;
constexpr double Mpc = 3.085678e22;
constexpr double kappa = 8. * pi * G;
constexpr double nu_const = 7. / 120 * pi * pi * pi * pi;
constexpr double nu_const2 = 5. / 7 / (pi * pi);
constexpr double zeta3 = 1.2020569031595942853997;
constexpr double zeta5 = 1.0369277551433699263313;
constexpr double am_min = 0.01, am_max = 600.;
constexpr double const_c = 2.99792458e8;  // source/settings.f90
}  // namespace bgc

// ---- host: massive-neutrino density table (built once per handle; constant across parameter points) ----
inline void build_nu_table(std::vector<double>& r1, std::vector<double>& dr1, double& dlnam) {
  const int n = NU_NRHOPN;
  std::vector<double> y(n + 1), g(n + 1), f(n + 1), dy(n + 1);
  dlnam = -(std::log(bgc::am_min / bgc::am_max)) / (n - 1);
  for (int i = 1; i <= n; i++) {
    const double am = bgc::am_min * std::exp((i - 1) * dlnam);
    // nuRhoPres: q up to 30 in 100 steps, spline-integrated, asymptotic tail
    const double qmax = 30.;
    const int nq = 100;
    double d1[nq + 2];
    const double adq = qmax / nq;
    d1[1] = 0;
    for (int k = 1; k <= nq; k++) {
      const double q = k * adq, aq = am / q;
      const double v = 1. / std::sqrt(1. + aq * aq);
      d1[k + 1] = adq * q * q * q / (std::exp(q) + 1.) / v;
    }
    const int m = nq + 1;
    const double dyn = (11. * d1[m] - 18. * d1[m - 1] + 9. * d1[m - 2] - 2. * d1[m - 3]) / 6.;
    double z = 0.5 * (d1[1] + d1[m]) + (0. - dyn) / 12.;
    double s = 0;
    for (int k = 2; k <= m - 1; k++) s += d1[k];
    z = z + s;
    y[i] = std::log((z + d1[m] / adq) / bgc::nu_const);
  }
  g[1] = 0;
  for (int i = 2; i <= n; i++) g[i] = 1 / (4. - g[i - 1]);
  f[1] = (-10. * y[1] + 15. * y[2] - 6. * y[3] + y[4]) / 6.;
  f[n] = (10. * y[n] - 15. * y[n - 1] + 6. * y[n - 2] - y[n - 3]) / 6.;
  for (int i = 2; i <= n - 1; i++) f[i] = g[i] * (3. * (y[i + 1] - y[i - 1]) - f[i - 1]);
  dy[n] = f[n];
  for (int i = n - 1; i >= 1; i--) dy[i] = f[i] - g[i] * dy[i + 1];
  r1.assign(y.begin() + 1, y.end());
  dr1.assign(dy.begin() + 1, dy.end());
}

struct BgTables {
  const double* r1;   // [2000] log rho_nu
  const double* dr1;  // [2000] d log rho_nu / d index
  double dlnam;
};

struct BgPoint {  // per-point densities (registers)
  double grhok, grhocb, grhog_nm, grhov, w_lam, r, omegab_h2_3e4;
  double grhormass[3], nu_masses[3];
  int n_eig, curv;  // curv: 0 flat, 1 closed, 2 open
};

__device__ __forceinline__ BgPoint bg_point(const double* __restrict__ bg) {
  BgPoint P;
  const double H0 = bg[0], omegab = bg[1], omegac = bg[2], omegan = bg[3], omegav = bg[4], tcmb = bg[6];
  P.w_lam = bg[5];
  P.n_eig = (int)bg[8];
  const double omegak = 1 - (omegab + omegac + omegav + omegan);
  P.curv = (fabs(omegak) <= 5e-7) ? 0 : (omegak < -5e-7 ? 1 : 2);
  if (P.curv == 0) P.r = 1;
  else {
    const double t = (bgc::c / 1000) / H0;
    P.r = 1. / sqrt(fabs(-omegak / (t * t)));
  }
  const double grhom = 3 * H0 * H0 / (bgc::c * bgc::c) * 1000 * 1000;
  const double grhog = bgc::kappa / (bgc::c * bgc::c) * 4 * bgc::sigma_boltz / (bgc::c * bgc::c * bgc::c) *
                       (tcmb * tcmb * tcmb * tcmb) * (bgc::Mpc * bgc::Mpc);
  const double grhor = 7. / 8 * pow(4. / 11, 4. / 3) * grhog;
  P.grhog_nm = grhog + grhor * bg[7];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const bool on = i < P.n_eig;
    P.grhormass[i] = on ? grhor * bg[9 + i] : 0.0;
    P.nu_masses[i] = on ? bgc::nu_const / (1.5 * bgc::zeta3) * grhom / grhor * omegan * bg[12 + i] / bg[9 + i] : 0.0;
  }
  P.grhocb = grhom * omegac + grhom * omegab;
  P.grhov = grhom * omegav;
  P.grhok = grhom * omegak;
  const double h = H0 / 100.0;
  P.omegab_h2_3e4 = 3.0e4 * omegab * (h * h);
  return P;
}

__device__ __forceinline__ double nu_rho(const BgTables& T, double am) {
  if (am <= bgc::am_min * 1.1) return 1. + bgc::nu_const2 * am * am;
  if (am >= bgc::am_max * 0.9) return 3 / (2 * bgc::nu_const) * (bgc::zeta3 * am + (15 * bgc::zeta5) / 2 / am);
  double d = log(am / bgc::am_min) / T.dlnam + 1.;
  const int i = (int)d;
  d = d - i;
  const double r0 = T.r1[i - 1], r1v = T.r1[i], d0 = T.dr1[i - 1], d1 = T.dr1[i];
  const double rhonu = r0 + d * (d0 + d * (3. * (r1v - r0) - 2. * d0 - d1 + d * (d0 + d1 + 2. * (r0 - r1v))));
  return exp(rhonu);
}

__device__ __forceinline__ double dtauda(const BgPoint& P, const BgTables& T, double a) {
  const double a2 = a * a;
  double grhoa2 = P.grhok * a2 + P.grhocb * a + P.grhog_nm;
  if (P.w_lam == -1.) grhoa2 = grhoa2 + P.grhov * a2 * a2;
  else grhoa2 = grhoa2 + P.grhov * pow(a, 1 - 3 * P.w_lam);
  for (int i = 0; i < P.n_eig; i++) grhoa2 = grhoa2 + nu_rho(T, a * P.nu_masses[i]) * P.grhormass[i];
  return sqrt(3 / grhoa2);
}

// mode 0: dtauda ; 1: dtauda*a (physical time) ; 2: dsound_da (approximate sound speed, CosmoMC theta)
template <int MODE>
__device__ __forceinline__ double bg_integrand(const BgPoint& P, const BgTables& T, double a) {
  const double d = dtauda(P, T, a);
  if (MODE == 1) return d * a;
  if (MODE == 2) {
    const double R = P.omegab_h2_3e4 * a;
    return d * (1.0 / sqrt(3 * (1 + R)));
  }
  return d;
}

template <int MODE>
__device__ double bg_rombint(const BgPoint& P, const BgTables& T, double a, double b, double tol) {
  const int MAXITER = 20, MAXJ = 5;
  double g[MAXJ + 2];
  double h = 0.5 * (b - a);
  double gmax = h * (bg_integrand<MODE>(P, T, a) + bg_integrand<MODE>(P, T, b));
  g[1] = gmax;
  int nint = 1;
  double error = 1.0e20, g0 = 0;
  int i = 0;
  for (;;) {
    i++;
    if (i > MAXITER || (i > 5 && fabs(error) < tol)) break;
    g0 = 0;
    for (int k = 1; k <= nint; k++) g0 = g0 + bg_integrand<MODE>(P, T, a + (k + k - 1) * h);
    g0 = 0.5 * g[1] + h * g0;
    h = 0.5 * h;
    nint = nint + nint;
    const int jmax = min(i, MAXJ);
    double fourj = 1;
#pragma unroll
    for (int j = 1; j <= MAXJ; j++) {
      if (j <= jmax) {
        fourj = 4 * fourj;
        const double g1 = g0 + (g0 - g[j]) / (fourj - 1);
        g[j] = g0;
        g0 = g1;
      }
    }
    if (fabs(g0) > tol) error = 1 - gmax / g0;
    else error = gmax;
    gmax = g0;
#pragma unroll
    for (int j = 1; j <= MAXJ + 1; j++)
      if (j == jmax + 1) g[j] = g0;
  }
  return g0;
}

__device__ __forceinline__ double bg_rofchi(const BgPoint& P, double chi) {
  if (P.curv == 1) return sin(chi);
  if (P.curv == 2) return sinh(chi);
  return chi;
}

__device__ __forceinline__ double bg_angular_diameter_distance(const BgPoint& P, const BgTables& T, double z) {
  const double chi = bg_rombint<0>(P, T, 1 / (1 + z), 1., 1e-4 / 1000);
  return P.r / (1 + z) * bg_rofchi(P, chi / P.r);
}

__device__ __forceinline__ double bg_hofz(const BgPoint& P, const BgTables& T, double z) {
  const double a = 1 / (1 + z);
  return 1 / (a * a * dtauda(P, T, a));
}

// D_A(z_j), H(z_j) for every point: thread = (z index, point), points fastest so a warp shares z
__global__ void __launch_bounds__(128) bg_distance_kernel(int np, int nz, const double* __restrict__ bg,
                                                          const double* __restrict__ z, BgTables T,
                                                          double* __restrict__ DA, double* __restrict__ H) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)np * nz) return;
  const int iz = (int)(t / np), pt = (int)(t % np);
  const BgPoint P = bg_point(bg + (size_t)pt * NBG);
  const double zz = z[iz];
  if (DA) DA[(size_t)pt * nz + iz] = bg_angular_diameter_distance(P, T, zz);
  if (H) H[(size_t)pt * nz + iz] = bg_hofz(P, T, zz);
}

// per point: tau0 (TimeOfz(0)), age in Gyr, CosmomcTheta   -> out[pt][3]
__global__ void __launch_bounds__(96) bg_scalars_kernel(int np, const double* __restrict__ bg, BgTables T,
                                                        double* __restrict__ out) {
  const int pt = blockIdx.x * 32 + (threadIdx.x & 31);
  const int what = threadIdx.x >> 5;  // warp 0: tau0, 1: age, 2: theta
  if (pt >= np) return;
  const BgPoint P = bg_point(bg + (size_t)pt * NBG);
  double v;
  if (what == 0) v = bg_rombint<0>(P, T, 0., 1., 1e-4 / 1000);
  else if (what == 1) v = bg_rombint<1>(P, T, 0., 1., 1e-4) * bgc::Mpc / bgc::c / bgc::Gyr;
  else {
    const double* b = bg + (size_t)pt * NBG;
    const double h = b[0] / 100.0;
    const double ombh2 = b[1] * (h * h), omdmh2 = (b[2] + b[3]) * (h * h);
    const double zstar = 1048 * (1 + 0.00124 * pow(ombh2, -0.738)) *
                         (1 + (0.0783 * pow(ombh2, -0.238) / (1 + 39.5 * pow(ombh2, 0.763))) *
                                  pow(omdmh2 + ombh2, 0.560 / (1 + 21.1 * pow(ombh2, 1.81))));
    const double astar = 1 / (1 + zstar);
    const double rs = bg_rombint<2>(P, T, 1e-8, astar, (double)1e-6f);
    const double DAv = bg_angular_diameter_distance(P, T, zstar) / astar;
    v = rs / DAv;
  }
  out[(size_t)pt * 3 + what] = v;
}

// ---- BAO (generic), MGS, HST: one thread per point ----------------------------------------------------
constexpr int BAO_MAXN = 16;
struct BaoParams {
  int np, nz_total, z_off, num_bao, kind;  // kind 0 generic quad-form, 1 MGS table
  int type[BAO_MAXN];
  double z[BAO_MAXN], obs[BAO_MAXN];
  const double* invcov;      // [num_bao][num_bao]
  const double* alpha_prob;  // MGS table
  int n_alpha;
  double rs_rescale, fixed_rs;
  const double* bg;  // [np][NBG]
  const double* DA;  // [np][nz_total]
  const double* H;
  double* out;
  int out_stride;
};

__global__ void bao_kernel(BaoParams p) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= p.np) return;
  const double* b = p.bg + (size_t)pt * NBG;
  const double rs_drag = p.fixed_rs > 0 ? p.fixed_rs : b[15];
  const double* DA = p.DA + (size_t)pt * p.nz_total + p.z_off;
  const double* H = p.H + (size_t)pt * p.nz_total + p.z_off;
  if (p.kind == 1) {
    const double z = p.z[0];
    const double ADD = DA[0] * (1. + z);
    const double Dv = pow(ADD * ADD * z / H[0], 1. / 3.);
    const double alphamgs = Dv / rs_drag / (638.9518 / 148.69);
    double r;
    if (alphamgs > 1.1985 || alphamgs < 0.8005) r = 1e30;
    else {
      const int ii = 1 + (int)floor((alphamgs - 0.8005) / (double)0.001f);
      r = (p.alpha_prob[ii - 1] + p.alpha_prob[ii]) / 2.0 / 2.0;
    }
    p.out[(size_t)pt * p.out_stride] = r;
    return;
  }
  const double rs = rs_drag * p.rs_rescale;
  double th[BAO_MAXN];
  for (int j = 0; j < p.num_bao; j++) {
    const double z = p.z[j];
    const double ADD = DA[j] * (1. + z);
    const double Dv = pow(ADD * ADD * z / H[j], 1. / 3.);
    double v = 0;
    switch (p.type[j]) {
      case 2: v = Dv / rs; break;
      case 7: v = bgc::const_c * H[j] / 1e3 * rs; break;
      case 8: v = bgc::const_c * H[j] / 1e3 * rs * 1.0e-3; break;
      case 3: v = rs / Dv; break;
      case 1: {
        const double omegak = 1 - (b[1] + b[2] + b[3] + b[4]);
        const double omegam = 1.0 - b[4] - omegak;
        const double hh = b[0] / 100;
        v = 100 * Dv * sqrt(omegam * hh * hh) / (bgc::const_c / 1e3 * z);
        break;
      }
      case 4: v = DA[j] / rs; break;
      case 10: v = (1 + z) * DA[j] / rs; break;
      case 5: v = (1 + z) * DA[j] * H[j]; break;
      default: v = nan(""); break;
    }
    th[j] = v - p.obs[j];
  }
  double s = 0;  // Matrix_QuadForm: v^T (M v)
  for (int i = 0; i < p.num_bao; i++) {
    double t = 0;
    for (int j = 0; j < p.num_bao; j++) t += p.invcov[i * p.num_bao + j] * th[j];
    s += th[i] * t;
  }
  p.out[(size_t)pt * p.out_stride] = s / 2;
}

__global__ void hst_kernel(int np, const double* __restrict__ bg, const double* __restrict__ DA, int nz_total,
                           int z_off, double zeff, double angconversion, double H0_obs, double H0_err,
                           double* __restrict__ out, int out_stride) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= np) return;
  const double th = zeff > 0 ? angconversion / DA[(size_t)pt * nz_total + z_off] : bg[(size_t)pt * NBG];
  out[(size_t)pt * out_stride] = (th - H0_obs) * (th - H0_obs) / (2 * H0_err * H0_err);
}

}  // namespace cb200
