// K7b — BICEP/Keck foreground model added to the map cross spectra before binning (BK15).
//
// Reference behaviour reproduced (source/CMB_BK_Planck.f90):
//   :109-147  DustScaling      greybody integral over the bandpass / pivot, band-centre error correction
//   :151-183  SyncScaling      power-law integral over the bandpass / pivot
//   :187-227  Decorrelation    Delta' = exp(log Delta * scl_nu * scl_ell)  (2 - exp(log(2-Delta)...) for Delta > 1)
//   :229-340  TBK_planck_AddForegrounds   dust / sync / correlated component added to EE and BB map spectra
// The foreground spectra are smooth closed forms in l, so they are never materialised per multipole in HBM: each
// CTA (one parameter point) builds the three l-shapes in shared memory and accumulates the binned foreground
// band powers  sum_l window[bin][pair][l] * FG_pair(l)  straight into the binned theory (the part the reference
// divides by the calibration, CMBlikes.f90:1108-1124).
#pragma once
#include "common.cuh"

namespace cb200 {

constexpr int BK_MAXMAPS = 24;
constexpr int BK_MAXL = 1024;

struct BkParams {
  int np, nmaps, nbins, ncl, lmin, lmax, n_nuis, nuis_off, lform_dust, lform_sync;
  int field[BK_MAXMAPS];     // 1 = E, 2 = B (0-based theory field: T,E,B,P)
  int bc_class[BK_MAXMAPS];  // band-centre error class: 0 none, 1: '95', 2: '150', 3: '220'
  int bp_off[BK_MAXMAPS + 1];
  double th_dust[BK_MAXMAPS], th_sync[BK_MAXMAPS], nu_bar[BK_MAXMAPS];
  double fpivot_dust, fpivot_sync, fp_dust_decorr[2], fp_sync_decorr[2];
  const double *bp_nu, *bp_R, *bp_dnu;
  const double *bp_lnnu, *bp_w;   // ln(nu) and R dnu per bandpass sample (host, at registration)
  const double* fgW;   // [lmax+1][nbins*ncl] (transposed at registration: coalesced over threads)
  const int* lrange;   // [nbins][2] multipole support of every bin's windows
  const double* nuis;  // [np][n_nuis]
  double* binned;      // [np][nbins*ncl]  (+=)
};

__device__ __forceinline__ double bk_decorr(double Delta, double nu0, double nu1, const double* nupivot, int l,
                                            int lform) {
  const double a = log(nu0 / nu1), b = log(nupivot[0] / nupivot[1]);
  const double scl_nu = (a * a) / (b * b);
  double scl_ell = 1.0;
  if (lform == 1) scl_ell = l / 80.0;
  else if (lform == 2) scl_ell = (l / 80.0) * (l / 80.0);
  if (Delta > 1) return 2.0 - exp(log(2.0 - Delta) * scl_nu * scl_ell);
  return exp(log(Delta) * scl_nu * scl_ell);
}

// frequency scalings of every map for one parameter point (DustScaling / SyncScaling incl. the band-centre error,
// CMB_BK_Planck.f90:109-183): CTA-wide sums over the bandpass samples; results in s_fd / s_fs / s_bc
__device__ __forceinline__ void bk_scalings(const BkParams& p, const double* d, double* s_fd, double* s_fs, double* s_bc,
                                            double (*s_red)[8]) {
  constexpr double T_CMB = 2.72548, hP = 6.62606957e-34, kB = 1.3806488e-23;
  constexpr double GK = hP / kB * 1e9;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double betadust = d[3], Tdust = d[4], betasync = d[6];
  for (int i = 0; i < p.nmaps; i++) {
    const int c = p.bc_class[i];
    const double bcerr = c == 0 ? 1.0 : d[12] + d[12 + c] + 1.;
    double gb = 0, pl = 0;
    for (int k = p.bp_off[i] + tid; k < p.bp_off[i + 1]; k += 256) {
      // nu^(3 + beta) as exp((3 + beta) ln nu): these two sums are ~8 000 samples per point and were the whole kernel
      const double nu = p.bp_nu[k], w = p.bp_w[k], ln = p.bp_lnnu[k];
      gb += w * exp((3 + betadust) * ln) / (exp(GK * nu / Tdust) - 1);
      pl += w * exp((2 + betasync) * ln);
    }
    gb = warp_sum(gb); pl = warp_sum(pl);
    __syncthreads();
    if (lane == 0) { s_red[0][warp] = gb; s_red[1][warp] = pl; }
    __syncthreads();
    if (tid == 0) {
      double gb_int = 0, pl_int = 0;
      for (int w = 0; w < 8; w++) { gb_int += s_red[0][w]; pl_int += s_red[1][w]; }
      const double nb = p.nu_bar[i];
      const double gb0 = pow(p.fpivot_dust, 3 + betadust) / (exp(GK * p.fpivot_dust / Tdust) - 1);
      const double pl0 = pow(p.fpivot_sync, 2 + betasync);
      double th_err = 1, gb_err = 1, pl_err = 1;
      if (bcerr != 1.) {
        const double e1 = exp(GK * nb / T_CMB) - 1, e2 = exp(GK * nb * bcerr / T_CMB) - 1;
        th_err = (bcerr * bcerr * bcerr * bcerr) * exp(GK * nb * (bcerr - 1) / T_CMB) * (e1 * e1) / (e2 * e2);
        gb_err = pow(bcerr, 3 + betadust) * (exp(GK * nb / Tdust) - 1) / (exp(GK * nb * bcerr / Tdust) - 1);
        pl_err = pow(bcerr, 2 + betasync);
      }
      s_fd[i] = (gb_int / gb0) / p.th_dust[i] * (gb_err / th_err);
      s_fs[i] = (pl_int / pl0) / p.th_sync[i] * (pl_err / th_err);
      s_bc[i] = bcerr;
    }
  }
}

__global__ void __launch_bounds__(256) bk_foreground_kernel(BkParams p) {
  __shared__ double s_fd[BK_MAXMAPS], s_fs[BK_MAXMAPS], s_bc[BK_MAXMAPS];
  __shared__ double s_dust[BK_MAXL], s_sync[BK_MAXL], s_ds[BK_MAXL];
  __shared__ double s_red[2][8];
  const int pt = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* d = p.nuis + (size_t)pt * p.n_nuis + p.nuis_off;
  const double Adust = d[0], Async = d[1], alphadust = d[2], betadust = d[3], Tdust = d[4], alphasync = d[5],
               betasync = d[6], dustsync_corr = d[7], EEtoBB_dust = d[8], EEtoBB_sync = d[9], Delta_dust = d[10],
               Delta_sync = d[11];
  bk_scalings(p, d, s_fd, s_fs, s_bc, s_red);
  for (int l = p.lmin + tid; l <= p.lmax; l += 256) {
    const double x = l / 80.0;
    s_dust[l] = Adust * pow(x, alphadust);
    s_sync[l] = Async * pow(x, alphasync);
    s_ds[l] = dustsync_corr * sqrt(Adust * Async) * pow(x, (alphadust + alphasync) / 2);
  }
  __syncthreads();
  const bool need_dust = fabs(Delta_dust - 1) > 1e-5, need_sync = fabs(Delta_sync - 1) > 1e-5;
  for (int t = tid; t < p.nbins * p.ncl; t += 256) {
    const int bin = t / p.ncl, c = t % p.ncl;
    // pair (i, j), i >= j, element index i(i+1)/2 + j
    int i = (int)((sqrt(8.0 * c + 1.0) - 1.0) / 2.0);
    while ((i + 1) * (i + 2) / 2 <= c) i++;
    while (i * (i + 1) / 2 > c) i--;
    const int j = c - i * (i + 1) / 2;
    const int fi = p.field[i], fj = p.field[j];
    if (!((fi == 1 && fj == 1) || (fi == 2 && fj == 2))) continue;
    double dust = s_fd[i] * s_fd[j], sync = s_fs[i] * s_fs[j], dustsync = s_fd[i] * s_fs[j] + s_fs[i] * s_fd[j];
    if (fi == 1) { dust *= EEtoBB_dust; sync *= EEtoBB_sync; dustsync *= sqrt(EEtoBB_dust * EEtoBB_sync); }
    const double nu_i = p.nu_bar[i] * s_bc[i], nu_j = p.nu_bar[j] * s_bc[j];
    const int NTc = p.nbins * p.ncl;
    const double* w = p.fgW + t;
    double acc = 0;
    // (a band-power window covers ~35 of the 600 multipoles: only the bin's support is walked)
    for (int l = p.lrange[2 * bin]; l <= p.lrange[2 * bin + 1]; l++) {
      const double wl = w[(size_t)l * NTc];
      if (wl == 0.0) continue;
      double dd = 1.0, dsy = 1.0;
      if (need_dust && i != j) dd = bk_decorr(Delta_dust, nu_i, nu_j, p.fp_dust_decorr, l, p.lform_dust);
      if (need_sync && i != j) dsy = bk_decorr(Delta_sync, nu_i, nu_j, p.fp_sync_decorr, l, p.lform_sync);
      acc += wl * (dust * s_dust[l] * dd + sync * s_sync[l] * dsy + dustsync * s_ds[l]);
    }
    p.binned[(size_t)pt * p.nbins * p.ncl + t] += acc;
  }
}

// ---- GEMM form of the same sum when no point of the batch asks for frequency decorrelation (Delta_dust = Delta_sync = 1):
// the foreground of a map pair is then  dust_ij x^alpha_d + sync_ij x^alpha_s + ds_ij x^((alpha_d+alpha_s)/2), three
// multipole shapes per parameter point, so the band powers of ALL pairs and bins are
//   G [3 np][nbins ncl] = Shapes [3 np][lmax+1] . W^T [lmax+1][nbins ncl]            (one DMMA GEMM for the batch)
// followed by a per-(point, pair) combination.  (ncu: the scalar kernel above spent 8 of the 34 ms of a 4 096-point BK15
// step walking 702 x 600 window entries per point, 30 instructions per multiply-add.)
__global__ void __launch_bounds__(256) bk_shapes_kernel(BkParams p, double* __restrict__ shapes /*[3 np][LW]*/,
                                                        double* __restrict__ scal /*[np][3][BK_MAXMAPS]*/) {
  __shared__ double s_fd[BK_MAXMAPS], s_fs[BK_MAXMAPS], s_bc[BK_MAXMAPS];
  __shared__ double s_red[2][8];
  const int pt = blockIdx.x, tid = threadIdx.x;
  const double* d = p.nuis + (size_t)pt * p.n_nuis + p.nuis_off;
  bk_scalings(p, d, s_fd, s_fs, s_bc, s_red);
  __syncthreads();
  if (tid < p.nmaps) {
    scal[((size_t)pt * 3 + 0) * BK_MAXMAPS + tid] = s_fd[tid];
    scal[((size_t)pt * 3 + 1) * BK_MAXMAPS + tid] = s_fs[tid];
    scal[((size_t)pt * 3 + 2) * BK_MAXMAPS + tid] = s_bc[tid];
  }
  const double Adust = d[0], Async = d[1], alphadust = d[2], alphasync = d[5], dustsync_corr = d[7];
  const int LW = p.lmax + 1;
  double* sh = shapes + (size_t)pt * 3 * LW;
  for (int l = tid; l < LW; l += 256) {
    double a = 0, b = 0, c = 0;
    if (l >= p.lmin) {
      const double lx = log(l / 80.0);
      a = Adust * exp(alphadust * lx);
      b = Async * exp(alphasync * lx);
      c = dustsync_corr * sqrt(Adust * Async) * exp((alphadust + alphasync) / 2 * lx);
    }
    sh[l] = a; sh[LW + l] = b; sh[2 * (size_t)LW + l] = c;
  }
}

__global__ void bk_combine_kernel(BkParams p, const double* __restrict__ G /*[3 np][nbins ncl]*/,
                                  const double* __restrict__ scal) {
  const int pt = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  const int NTc = p.nbins * p.ncl;
  if (pt >= p.np || t >= NTc) return;
  const int c = t % p.ncl;
  int i = (int)((sqrt(8.0 * c + 1.0) - 1.0) / 2.0);
  while ((i + 1) * (i + 2) / 2 <= c) i++;
  while (i * (i + 1) / 2 > c) i--;
  const int j = c - i * (i + 1) / 2;
  const int fi = p.field[i], fj = p.field[j];
  if (!((fi == 1 && fj == 1) || (fi == 2 && fj == 2))) return;
  const double* d = p.nuis + (size_t)pt * p.n_nuis + p.nuis_off;
  const double EEtoBB_dust = d[8], EEtoBB_sync = d[9];
  const double* fd = scal + ((size_t)pt * 3 + 0) * BK_MAXMAPS;
  const double* fs = scal + ((size_t)pt * 3 + 1) * BK_MAXMAPS;
  double dust = fd[i] * fd[j], sync = fs[i] * fs[j], dustsync = fd[i] * fs[j] + fs[i] * fd[j];
  if (fi == 1) { dust *= EEtoBB_dust; sync *= EEtoBB_sync; dustsync *= sqrt(EEtoBB_dust * EEtoBB_sync); }
  const double* g = G + (size_t)pt * 3 * NTc + t;
  p.binned[(size_t)pt * NTc + t] += dust * g[0] + sync * g[NTc] + dustsync * g[2 * (size_t)NTc];
}

}  // namespace cb200
