// Gaussian / quadratic-form CMB likelihood kernels (K7, K8 of DESIGN.md).
//
// Reference behaviour reproduced (paths relative to the reference root):
//   source/CMB.f90:305-329       TPlikLiteLikelihood_LogLike  (band powers, calibration, chi^2)
//   source/CMBlikes.f90:1165-1256 CMBLikes_LogLike  (binned; gaussian and Hamimeche-Lewis branches)
//   source/CMBlikes.f90:861-914   CMBLikes_Transform
//   source/Matrix_utils_new.f90:2033-2047 Matrix_QuadForm
// chi^2 = x^T C^-1 x is evaluated for the whole batch as one FP64 tensor-pipe GEMM  T = X * C^-1  followed by
// a row-wise dot product.
#pragma once
#include "common.cuh"

namespace cb200 {

// plik-lite band powers and residuals: thread per (point, used bin)
struct PlikBinParams {
  int np, nused, lmax_out, n_nuis, cal_index;
  const double* cls;     // [np][5][lmax_out+1] TT,TE,EE,BB,PP
  const int* bin_spec;   // [nused] 0 TT, 1 TE, 2 EE
  const int* bin_lo;     // [nused] absolute l
  const int* bin_hi;     // [nused]
  const double* weights; // [lmax_w+1] by l (already * 2pi/(l(l+1)))
  const double* x_data;  // [nused]
  const double* nuis;    // [np][n_nuis]
  double* resid;         // [np][nused]
};

__global__ void plik_bin_kernel(PlikBinParams p) {
  const int lp = blockIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= p.np || b >= p.nused) return;
  const double* c = p.cls + ((size_t)lp * 5 + p.bin_spec[b]) * (p.lmax_out + 1);
  double s = 0;
  for (int l = p.bin_lo[b]; l <= p.bin_hi[b]; l++) s += c[l] * p.weights[l];
  const double cal = (p.cal_index >= 0) ? p.nuis[(size_t)lp * p.n_nuis + p.cal_index] : 1.0;
  p.resid[(size_t)lp * p.nused + b] = p.x_data[b] - s / (cal * cal);
}

// row-wise dot: out[lp] (+)= scale * sum_i T[lp][i] * X[lp][i]  ; one warp per point
__global__ void rowdot_kernel(int np, int n, const double* __restrict__ T, const double* __restrict__ X,
                              double scale, double* __restrict__ out, int out_stride, int accumulate) {
  const int lp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (lp >= np) return;
  double s = 0;
  for (int i = lane; i < n; i += 32) s += T[(size_t)lp * n + i] * X[(size_t)lp * n + i];
  s = warp_sum(s);
  if (lane == 0) {
    double v = scale * s;
    out[(size_t)lp * out_stride] = accumulate ? out[(size_t)lp * out_stride] + v : v;
  }
}

// ---- generic binned CMBLikes ---------------------------------------------------------------------------
constexpr int CMBL_MAXMAPS = 12;

// symmetric eigen-decomposition by cyclic Jacobi (replaces LAPACK DSYEV of Matrix_Diagonalize,
// source/Matrix_utils_new.f90:361-383); eigenvalues unsorted, A destroyed, V columns = eigenvectors.
__device__ void jacobi_eigen(double* A, double* V, double* w, int n) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) {
      diag += A[i * n + i] * A[i * n + i];
      for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    }
    if (off <= 1e-32 * diag || off == 0) break;
    for (int pp = 0; pp < n; pp++)
      for (int q = pp + 1; q < n; q++) {
        const double apq = A[pp * n + q];
        if (apq == 0) continue;
        const double tau = (A[q * n + q] - A[pp * n + pp]) / (2 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1 + tau * tau));
        const double c = 1 / sqrt(1 + t * t), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + pp], akq = A[k * n + q];
          A[k * n + pp] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[pp * n + k], aqk = A[q * n + k];
          A[pp * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + pp], vkq = V[k * n + q];
          V[k * n + pp] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

struct CmbLikesBinParams {
  int np, nmaps, ncl, nbins, ncl_used, like_approx, n_nuis, cal_index;
  const double* bc;      // [np][nbins*ncl] CMB-spectra part of the binned theory (to be / cal^2)
  const double* bp;      // [np][nbins*ncl] lensing-potential part
  const double* offset;  // [nbins*ncl]
  const double* noise;   // [nbins][nmaps*nmaps] or null
  const double* chat;    // [nbins][nmaps*nmaps]
  const double* sqrt_fid;  // [nbins][nmaps*nmaps] (HL)
  const int* cl_use;     // [ncl_used]
  const double* nuis;
  double* bigx;          // [np][nbins*ncl_used]
  double* binned_out;    // optional [np][nbins*ncl] (after correction; for parity tests)
};

// thread per (point, bin), the pairs packed densely into warps: with the 9 bins of a data set spread over a 32-thread
// block per point, 23 of 32 lanes idled through two 12 x 12 Jacobi eigen-solves (30 of the 43 ms of a 4 096-point BK15 step)
constexpr int CMBL_THREADS = 128;
__global__ void __launch_bounds__(CMBL_THREADS) cmblikes_bin_kernel(CmbLikesBinParams p) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)p.np * p.nbins) return;
  const int lp = (int)(t / p.nbins);
  const int bin = (int)(t - (long long)lp * p.nbins);
  const int n = p.nmaps;
  const double cal = (p.cal_index >= 0) ? p.nuis[(size_t)lp * p.n_nuis + p.cal_index] : 1.0;
  const double ic2 = 1.0 / (cal * cal);
  double C[CMBL_MAXMAPS * CMBL_MAXMAPS];
  double vecp[CMBL_MAXMAPS * (CMBL_MAXMAPS + 1) / 2];
  int ix = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      const size_t o = (size_t)bin * p.ncl + ix;
      const double v = p.bc[(size_t)lp * p.nbins * p.ncl + o] * ic2 + p.bp[(size_t)lp * p.nbins * p.ncl + o] - p.offset[o];
      if (p.binned_out) p.binned_out[(size_t)lp * p.nbins * p.ncl + o] = v;
      C[i * n + j] = v;
      C[j * n + i] = v;
      ix++;
    }
  if (p.noise)
    for (int k = 0; k < n * n; k++) C[k] += p.noise[(size_t)bin * n * n + k];
  const double* Chat = p.chat + (size_t)bin * n * n;
  if (p.like_approx == 1) {
    // Hamimeche-Lewis: C <- Cf^1/2 U g(D) U^T Cf^1/2 with C^-1/2 Chat C^-1/2 = U D U^T
    const double* Cf = p.sqrt_fid + (size_t)bin * n * n;
    double U[CMBL_MAXMAPS * CMBL_MAXMAPS], R[CMBL_MAXMAPS * CMBL_MAXMAPS], T[CMBL_MAXMAPS * CMBL_MAXMAPS];
    double d[CMBL_MAXMAPS];
    jacobi_eigen(C, U, d, n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += U[k * n + i] * Chat[k * n + j];
        T[i * n + j] = s;
      }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += T[i * n + k] * U[k * n + j];
        R[i * n + j] = s / (sqrt(d[i]) * sqrt(d[j]));
      }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += R[i * n + k] * U[j * n + k];
        T[i * n + j] = s;
      }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += U[i * n + k] * T[k * n + j];
        R[i * n + j] = s;
      }
    jacobi_eigen(R, U, d, n);
    for (int i = 0; i < n; i++) {
      const double v = sqrt(2 * fmax(0.0, d[i] - log(d[i]) - 1));
      d[i] = (d[i] - 1 >= 0) ? v : -v;
    }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += Cf[i * n + k] * U[k * n + j];
        T[i * n + j] = s;
      }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += T[i * n + k] * d[k] * T[j * n + k];
        C[i * n + j] = s;
      }
  } else {
    for (int k = 0; k < n * n; k++) C[k] -= Chat[k];
  }
  ix = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) vecp[ix++] = C[i * n + j];
  for (int k = 0; k < p.ncl_used; k++)
    p.bigx[(size_t)lp * p.nbins * p.ncl_used + (size_t)bin * p.ncl_used + k] = vecp[p.cl_use[k]];
}

// -lnL = chi2/2 (+ calibration prior): out[lp] = 0.5*(q[lp] + (ln cal/prior)^2)
__global__ void cmblikes_final_kernel(int np, const double* __restrict__ quad, const double* __restrict__ nuis,
                                      int n_nuis, int cal_index, double log_cal_prior, double* __restrict__ out,
                                      int out_stride) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np) return;
  double chisq = quad[lp];
  if (log_cal_prior > 0 && cal_index >= 0) {
    const double t = log(nuis[(size_t)lp * n_nuis + cal_index]) / log_cal_prior;
    chisq += t * t;
  }
  out[(size_t)lp * out_stride] = chisq / 2;
}

__global__ void total_kernel(int np, int n_like, const double* __restrict__ ll, const int* __restrict__ status,
                             double* __restrict__ total) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= np) return;
  double s = 0;
  for (int i = 0; i < n_like; i++) s += ll[(size_t)lp * n_like + i];
  if (status && status[lp] != 0) s = 1e30;
  if (isnan(s)) s = 1e30;
  total[lp] = s;
}

}  // namespace cb200
