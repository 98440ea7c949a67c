// Line-of-sight projection of the source functions against the spherical-Bessel table, fused with the
// P(k)-weighted k-contraction (kernels K0, K1, K2 of DESIGN.md).
//
// Reference behaviour reproduced (paths relative to the reference root):
//   K0  camb/cmbmain.f90:1207-1218  InitSourceInterpolation  (natural spline of Src along k)
//   K1  camb/cmbmain.f90:478-498    SourceToTransfers
//       camb/cmbmain.f90:1295-1374  InterpolateSources
//       camb/cmbmain.f90:1387-1420  DoSourceIntegration (flat)
//       camb/cmbmain.f90:1440-1562  DoFlatIntegration incl. the Limber switch for the lensing source
//   K2  camb/cmbmain.f90:2132-2264  CalcScalCls, :2344-2397 CalcTensCls (flat)
//
// Parallel decomposition (B200): one CTA per (parameter point, block of Q consecutive wavenumbers).
// Lanes of a warp are consecutive sampled multipoles, so the Bessel node table (multipole fastest) is read
// 512 contiguous bytes per node; NS warp groups split the conformal-time axis and are reduced through
// shared memory; the transfer functions Delta_l(q) never leave registers: the CTA emits its partial
// sum over q of P(q) Delta_a Delta_b dlnq, and a small second kernel adds the partials in a fixed order.
#pragma once
#include "common.cuh"

namespace cb200 {

constexpr int PROJ_Q = 8;      // wavenumbers per CTA
constexpr int PROJ_NS = 4;     // warp groups splitting the time axis
constexpr int PROJ_SLAB = 32;  // time samples per group per pass
constexpr int PROJ_LW = 3;     // multipole warps (96 padded multipoles)
constexpr int PROJ_LP = 32 * PROJ_LW;

struct PointView {  // resident per-point inputs (device pointers), strides NT / NK / NQ
  int NT, NK, NQ, NSRC;
  const double* thermo;  // [P][5] tau0, taurst, taurend, reion_start, reion_complete
  const int* n_tau;
  const double* tau;     // [P][NT]
  const double* dtau;    // [P][NT]
  const LinSegs* tseg;   // [P]
  const int* n_k;
  const double* ksrc;    // [P][NK]
  const int* n_q;
  const double* q;       // [P][NQ]
  const double* dq;      // [P][NQ]
  const double* src;     // [P][NT][NSRC][NK]
};

// ------------------------------------------------------------------------------------------------ K0
// Spline set-up that depends on the k grid only (per point): xxdiv, sig, xp, gam of the tridiagonal sweep.
__global__ void spline_setup_kernel(PointView v, int p0, int np, double* __restrict__ coef /*[np][5][NK]*/) {
  int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np) return;
  const int pt = p0 + lp, nk = v.n_k[pt];
  const double* x = v.ksrc + (size_t)pt * v.NK;
  double* c = coef + (size_t)lp * 5 * v.NK;
  double gam = 0;
  c[3 * v.NK + 0] = 0;
  for (int i = 0; i + 1 < nk; i++) c[4 * v.NK + i] = 1. / (x[i + 1] - x[i]);  // used by the tiled kernel only
  for (int i = 1; i <= nk - 2; i++) {
    double xxdiv = 1. / (x[i + 1] - x[i - 1]);
    double sig = (x[i] - x[i - 1]) * xxdiv;
    double xp = 1. / (sig * gam + 2.);
    gam = (sig - 1.) * xp;
    c[0 * v.NK + i] = xxdiv;
    c[1 * v.NK + i] = sig;
    c[2 * v.NK + i] = xp;
    c[3 * v.NK + i] = gam;
  }
}

// one thread per (point, tau, source) row: forward sweep for u, backward sweep for the second derivatives
__global__ void source_spline_kernel(PointView v, int p0, int np, const double* __restrict__ coef,
                                     double* __restrict__ ddsrc /*[np][NT][NSRC][NK]*/) {
  const long long rows_per_pt = (long long)v.NT * v.NSRC;
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows_per_pt * np) return;
  const int lp = (int)(r / rows_per_pt);
  const int row = (int)(r % rows_per_pt);
  const int pt = p0 + lp;
  const int it = row / v.NSRC;
  if (it >= v.n_tau[pt]) return;
  const int nk = v.n_k[pt];
  const double* x = v.ksrc + (size_t)pt * v.NK;
  const double* y = v.src + ((size_t)pt * rows_per_pt + row) * v.NK;
  double* d2 = ddsrc + ((size_t)lp * rows_per_pt + row) * v.NK;
  const double* c = coef + (size_t)lp * 5 * v.NK;
  double d1r = (y[1] - y[0]) / (x[1] - x[0]), d1l, u = 0;
  d2[0] = 0;
  for (int i = 1; i <= nk - 2; i++) {
    d1l = d1r;
    d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    u = (6. * (d1r - d1l) * c[0 * v.NK + i] - c[1 * v.NK + i] * u) * c[2 * v.NK + i];
    d2[i] = u;
  }
  double nxt = 0;
  d2[nk - 1] = 0;
  for (int i = nk - 2; i >= 1; i--) {
    nxt = c[3 * v.NK + i] * nxt + d2[i];
    d2[i] = nxt;
  }
  d2[0] = 0;  // gam(0) = 0, u(0) = 0
}

// Tiled K0: the rows of one point are contiguous ([tau][source][k]), so a CTA stages SPL_ROWS consecutive rows in
// shared memory with coalesced 16-byte loads, lets one thread per row run the two sequential sweeps out of shared
// memory (row stride NK+1 words: conflict-free), and writes the second derivatives back coalesced.  The one-thread-
// per-row kernel above reads and writes global memory with a 1.8 KB stride between lanes (0.65 TB/s measured).
// 1/(x[i+1]-x[i]) comes from the set-up kernel (a multiply instead of the reference's division: values, not indices).
constexpr int SPL_ROWS = 32;
constexpr int SPL_THREADS = 128;
__global__ void __launch_bounds__(SPL_THREADS) source_spline_tiled_kernel(PointView v, int p0, int np,
                                                                           const double* __restrict__ coef,
                                                                           double* __restrict__ ddsrc) {
  extern __shared__ __align__(16) double spl_smem[];
  const int NK = v.NK, NKP = NK + 1;
  double* sc = spl_smem;                 // [5][NK] sweep coefficients + reciprocal spacings
  double* sy = spl_smem + 5 * NK;        // [SPL_ROWS][NKP]
  const int rows_per_pt = v.NT * v.NSRC;
  const int tiles_per_pt = (rows_per_pt + SPL_ROWS - 1) / SPL_ROWS;
  const int lp = blockIdx.x / tiles_per_pt, tile = blockIdx.x - lp * tiles_per_pt;
  if (lp >= np) return;
  const int pt = p0 + lp;
  const int row0 = tile * SPL_ROWS;
  const int nrow_pt = v.n_tau[pt] * v.NSRC;          // rows beyond the point's time samples are never read
  if (row0 >= nrow_pt) return;
  const int nrows = min(SPL_ROWS, nrow_pt - row0);
  const int nk = v.n_k[pt];
  const double* c = coef + (size_t)lp * 5 * NK;
  const double* y = v.src + ((size_t)pt * rows_per_pt + row0) * NK;
  double* d2 = ddsrc + ((size_t)lp * rows_per_pt + row0) * NK;
  const int tid = threadIdx.x;
  // sweep coefficients.  The reference's forward step u_i = (6 (d_i - d_i-1) c0_i - c1_i u_i-1) c2_i puts three dependent
  // FP64 operations per wavenumber on ONE thread's chain (FP64 latency, not bandwidth, is what this kernel waits for);
  // with c1 c2 and 6 c0 c2 formed here the chain is one fused multiply-add per step (rounding differs in the last bit)
  for (int e = tid; e < NK; e += SPL_THREADS) {
    const double c0 = c[e], c1 = c[NK + e], c2 = c[2 * NK + e];
    sc[e] = 6. * c0 * c2;
    sc[NK + e] = c1 * c2;
    sc[2 * NK + e] = c2;
    sc[3 * NK + e] = c[3 * NK + e];
    sc[4 * NK + e] = c[4 * NK + e];
  }
  // NK is a multiple of 2: 16-byte loads of the contiguous block, scattered into the padded rows
  const double2* y2 = reinterpret_cast<const double2*>(y);
  const int n2 = nrows * NK / 2;
  // eight loads in flight per thread: one load per iteration left the 28 iterations of this loop at one DRAM latency each
  constexpr int SPL_U = 8;
  for (int e0 = tid; e0 < n2; e0 += SPL_U * SPL_THREADS) {
    double2 val[SPL_U];
#pragma unroll
    for (int u = 0; u < SPL_U; u++) {
      const int e = e0 + u * SPL_THREADS;
      if (e < n2) val[u] = __ldg(y2 + e);
    }
#pragma unroll
    for (int u = 0; u < SPL_U; u++) {
      const int e = e0 + u * SPL_THREADS;
      if (e < n2) {
        const int r = (2 * e) / NK, k = 2 * e - r * NK;
        sy[r * NKP + k] = val[u].x;
        sy[r * NKP + k + 1] = val[u].y;
      }
    }
  }
  __syncthreads();
  if (tid < nrows) {
    double* yr = sy + tid * NKP;
    double y0 = yr[0], y1 = yr[1];
    double d1r = (y1 - y0) * sc[4 * NK + 0], d1l, u = 0;
    yr[0] = 0;
    for (int i = 1; i <= nk - 2; i++) {
      const double y2v = yr[i + 1];
      d1l = d1r;
      d1r = (y2v - y1) * sc[4 * NK + i];
      u = fma(-sc[1 * NK + i], u, (d1r - d1l) * sc[0 * NK + i]);
      yr[i] = u;          // y[i] is no longer needed
      y1 = y2v;
    }
    double nxt = 0;
    yr[nk - 1] = 0;
    for (int i = nk - 2; i >= 1; i--) {
      nxt = sc[3 * NK + i] * nxt + yr[i];
      yr[i] = nxt;
    }
    for (int i = nk; i < NK; i++) yr[i] = 0;
  }
  __syncthreads();
  double2* d22 = reinterpret_cast<double2*>(d2);
  for (int e = tid; e < n2; e += SPL_THREADS) {
    const int r = (2 * e) / NK, k = 2 * e - r * NK;
    d22[e] = make_double2(sy[r * NKP + k], sy[r * NKP + k + 1]);
  }
}

// ---- primordial power spectra, shared by every projection kernel and by the shared-transfer contraction
__device__ __forceinline__ double scalar_power_dev(const double* ip, double k) {
  // camb/power_tilt.f90:114-133
  double lnrat = log(k / ip[7]);
  return ip[0] * exp(lnrat * (ip[1] - 1 + lnrat * (ip[2] / 2 + ip[3] / 6 * lnrat)));
}
__device__ __forceinline__ double tensor_power_dev(const double* ip, double k) {
  // camb/power_tilt.f90:136-162 with the CosmoMC mapping source/Calculator_CAMB.f90:839-877
  double r = ip[4], ns = ip[1], nt, ntrun;
  if (ip[9] != 0) {
    nt = -r / 8 * (2 - ns - r / 8);
    ntrun = r / 8 * (r / 8 + ns - 1);
  } else { nt = ip[5]; ntrun = ip[6]; }
  double lnrat = log(k / ip[8]);
  double k_dep = exp(lnrat * (nt + ntrun / 2 * lnrat));
  if (ip[8] != ip[7]) return r * scalar_power_dev(ip, ip[8]) * k_dep;  // tensor_param_rpivot
  return r * ip[0] * k_dep;                                            // tensor_param_indeptilt
}

// ------------------------------------------------------------------------------------------------ K2
// fixed-order sum of the per-CTA partials + the l-dependent normalisations (cmbmain.f90:2222-2257, :2388-2393)
__global__ void contract_reduce_kernel(int np, int p0, const int* __restrict__ n_q, int Q, int NQB, int nl,
                                       const int* __restrict__ ls, int tensors, const double* __restrict__ alens,
                                       const double* __restrict__ part, double* __restrict__ icl /*[np][6][PROJ_LP]*/,
                                       int nq_shared /* > 0: every point uses this wavenumber count */) {
  int lp = blockIdx.y;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np || t >= 6 * PROJ_LP) return;
  const int X = t / PROJ_LP, j = t % PROJ_LP;
  double s = 0;
  if (j < nl) {
    const int nb = ((nq_shared > 0 ? nq_shared : n_q[p0 + lp]) + Q - 1) / Q;
    const double* pp = part + (((size_t)lp * NQB) * 6 + X) * PROJ_LP + j;
    for (int b = 0; b < nb; b++) s += pp[(size_t)b * 6 * PROJ_LP];
    const double ell = ls[j];
    const double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
    if (!tensors) {
      const double dbletmp = (ell * (ell + 1)) / kTwoPi * kFourPi;
      const double AL = alens ? alens[lp] : 1.0;
      switch (X) {
        case 0: s = s * dbletmp; break;
        case 1: s = s * dbletmp * ctnorm; break;
        case 2: s = s * dbletmp * sqrt(ctnorm); break;
        case 3: s = AL * s * kFourPi * ell * ell * ell * ell; break;
        case 4: s = sqrt(AL) * s * kFourPi * ell * ell * ell; break;
        default: s = sqrt(AL) * s * kFourPi * ell * ell * ell * sqrt(ctnorm); break;
      }
    } else {
      double dbletmp = (ell * (ell + 1)) / kTwoPi * kPi / 4;
      switch (X) {
        case 0: s = s * dbletmp * ctnorm; break;
        case 1: case 2: s = (ls[j] == 1) ? 0.0 : s * dbletmp; break;
        case 3: s = s * dbletmp * sqrt(ctnorm); break;
        default: s = 0; break;
      }
    }
  }
  icl[((size_t)lp * 6 + X) * PROJ_LP + j] = s;
}

// ---- shared-transfer contraction (one source point, many initial-power points): operands of the DMMA GEMM ----
// D2[q][X][l] = Delta_a Delta_b of CalcScalCls / CalcTensCls (cmbmain.f90:2195-2215, :2377-2383)
__global__ void delta_products_kernel(int nq, int tensors, const double* __restrict__ delta /*[nq][PROJ_LP][3]*/,
                                      double* __restrict__ d2 /*[nq][6][PROJ_LP]*/) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * PROJ_LP) return;
  const int q = t / PROJ_LP, j = t % PROJ_LP;
  const double* d = delta + (size_t)t * 3;
  const double d0 = d[0], d1 = d[1], dd2 = d[2];
  double* o = d2 + (size_t)q * 6 * PROJ_LP + j;
  if (tensors) {
    o[0] = d0 * d0; o[PROJ_LP] = d1 * d1; o[2 * PROJ_LP] = dd2 * dd2; o[3 * PROJ_LP] = d0 * d1;
    o[4 * PROJ_LP] = 0; o[5 * PROJ_LP] = 0;
  } else {
    o[0] = d0 * d0; o[PROJ_LP] = d1 * d1; o[2 * PROJ_LP] = d0 * d1;
    o[3 * PROJ_LP] = dd2 * dd2; o[4 * PROJ_LP] = dd2 * d0; o[5 * PROJ_LP] = dd2 * d1;
  }
}
// W[pt][q] = P(q; initpower_pt) dq/q
__global__ void power_weights_kernel(int np, int nq, int tensors, const double* __restrict__ q,
                                     const double* __restrict__ dq, const double* __restrict__ initpower,
                                     double* __restrict__ W, int ldw) {
  const int lp = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np || i >= nq) return;
  const double* ip = initpower + (size_t)lp * 10;
  const double qv = q[i];
  W[(size_t)lp * ldw + i] = (tensors ? tensor_power_dev(ip, qv) : scalar_power_dev(ip, qv)) * (dq[i] / qv);
}

}  // namespace cb200
