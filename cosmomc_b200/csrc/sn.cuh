// K6 — JLA / Pantheon supernova likelihood, batched over parameter points.
//
// Reference behaviour reproduced (source/supernovae_JLA.f90):
//   :874-991    jla_prep            pre_vars, A1/A2 masks of the two-scriptM fit
//   :773-866    invert_covariance_matrix   V(alpha,beta) = V0 + a^2 Va + b^2 Vb + 2a V0a - 2b V0b - 2ab Vab + diag
//   :1028-1168  JLA_alpha_beta_like        scriptM pre-estimate, A..F marginalisation terms, chi^2
//   :1170-1228  jla_LnLike                 lumdists = 5 log10((1+zhel)(1+zcmb) D_A(zcmb))
// The reference factors V (DPOTRF), forms the full inverse (DPOTRI) and multiplies (DSYMV).  Here V = L L^T is
// factored by a blocked left-looking Cholesky on the FP64 tensor pipe (DMMA m8n8k4); V(alpha, beta) is assembled from
// the six constant L2-resident blocks by sn_assemble_kernel ahead of the factorisation (one coalesced pass; the
// in-kernel assembly of the first version put six dependent loads per element into the panel's latency chain and is
// kept only behind CholParams::assemble), and the right-hand sides [d, A1, A2] ride
// along as extra rows of the panel, so that after the factorisation those rows hold y = L^-1 rhs and
//   A = y_d.y_d, B = y_d.y_1, C = y_d.y_2, D = y_1.y_2, E = y_1.y_1, F = y_2.y_2
// (algebraically identical to the V^-1 expressions, without ever forming V^-1).
#pragma once
#include "common.cuh"
#include "dgemm.cuh"

namespace cb200 {

struct SnData {  // device pointers, all [nsn]
  int nsn;
  const double *zcmb, *zhel, *mag, *stretch, *colour, *pre_vars, *stretch_var, *colour_var, *cov_ms, *cov_mc, *cov_sc;
  const double *A1, *A2;  // A1 = ones when !twoscriptm
  const double* cov[6];   // mag, stretch, colour, mag_stretch, mag_colour, stretch_colour (null if absent), [nsn][nsn]
  int twoscriptm, alphabeta;
};

__device__ __forceinline__ double sn_diag(const SnData& S, int i, double alpha, double beta) {
  return S.pre_vars[i] + alpha * alpha * S.stretch_var[i] + beta * beta * S.colour_var[i] + 2.0 * alpha * S.cov_ms[i] -
         2.0 * beta * S.cov_mc[i] - 2.0 * alpha * beta * S.cov_sc[i];
}

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sh[w];
  return s;
}

// per point: luminosity distances, scriptM pre-estimate, diffmag -> rhs rows of the work matrix (or a dense [np][nsn])
//   DA [np][nz_total] (+z_off) ; ab [np][n_nuis] with alpha at ia, beta at ib (<0: 0)
__global__ void __launch_bounds__(256) sn_prep_kernel(SnData S, int np, const double* __restrict__ DA, int nz_total,
                                                      int z_off, const double* __restrict__ nuis, int n_nuis, int ia,
                                                      int ib, double* __restrict__ rhs, size_t rhs_pt_stride,
                                                      int rhs_row_stride, int write_masks) {
  __shared__ double sh[8];
  const int pt = blockIdx.x;
  const double alpha = ia >= 0 ? nuis[(size_t)pt * n_nuis + ia] : 0.0;
  const double beta = ib >= 0 ? nuis[(size_t)pt * n_nuis + ib] : 0.0;
  const double* da = DA + (size_t)pt * nz_total + z_off;
  double w = 0, ws = 0;
  for (int i = threadIdx.x; i < S.nsn; i += blockDim.x) {
    const double lum = 5.0 * log10((1.0 + S.zhel[i]) * (1.0 + S.zcmb[i]) * da[i]);
    const double iv = 1.0 / sn_diag(S, i, alpha, beta);
    w += iv;
    ws += (S.mag[i] - lum) * iv;
  }
  const double wtval = block_sum_256(w, sh);
  const double est = block_sum_256(ws, sh) / wtval;
  double* r = rhs + (size_t)pt * rhs_pt_stride;
  for (int i = threadIdx.x; i < S.nsn; i += blockDim.x) {
    const double lum = 5.0 * log10((1.0 + S.zhel[i]) * (1.0 + S.zcmb[i]) * da[i]);
    r[i] = S.mag[i] - lum + alpha * S.stretch[i] - beta * S.colour[i] - est;
    if (write_masks) {
      r[(size_t)rhs_row_stride + i] = S.A1[i];
      if (S.twoscriptm) r[(size_t)2 * rhs_row_stride + i] = S.A2[i];
    }
  }
}

// ---- V(alpha, beta) of every point, lower triangle, coalesced (supernovae_JLA.f90:1074-1096) -----------------------
// One warp per matrix row, lanes over the columns: the six constant blocks (26 MB, L2-resident across the batch) are
// read 256 B at a time and the row goes to W[pt][row][0..row].  Doing this ahead of the factorisation takes the six
// dependent, poorly coalesced loads per element out of the Cholesky's latency chain (ncu: 15 % of its samples).
__global__ void __launch_bounds__(256) sn_assemble_kernel(SnData S, int np, const double* __restrict__ nuis, int n_nuis,
                                                          int ia, int ib, double* __restrict__ W, size_t pt_stride, int ld) {
  const int pt = blockIdx.y, row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int n = S.nsn;
  if (pt >= np || row >= n) return;
  const double alpha = ia >= 0 ? nuis[(size_t)pt * n_nuis + ia] : 0.0;
  const double beta = ib >= 0 ? nuis[(size_t)pt * n_nuis + ib] : 0.0;
  const double cf[6] = {1.0, alpha * alpha, beta * beta, 2.0 * alpha, -2.0 * beta, -2.0 * alpha * beta};
  double* out = W + (size_t)pt * pt_stride + (size_t)row * ld;
  const int cmax = min(n - 1, row | 31);  // through the end of the row's diagonal block (the factorisation loads whole blocks)
  for (int col = lane; col <= cmax; col += 32) {
    const size_t o = (size_t)row * n + col;
    double av = 0.0;
#pragma unroll
    for (int m = 0; m < 6; m++)
      if (S.cov[m]) av += cf[m] * S.cov[m][o];
    if (row == col) av += sn_diag(S, row, alpha, beta);
    out[col] = av;
  }
}

// ---- blocked Cholesky with ride-along right-hand sides ----------------------------------------------------
// W[pt]: (n + nr) rows x ld, row-major.  Rows 0..n-1: lower triangle of V (assembled on the fly when S != null,
// else read from W), rows n..n+nr-1: right-hand sides.  On exit rows 0..n-1 hold L (lower), rows n.. hold (L^-1 rhs)^T.
constexpr int CH_NB = 32, CH_ROWS = 128, CH_KC = 32, CH_AS = CH_KC + 4, CH_CS = CH_NB + 1;
constexpr size_t CH_SMEM = sizeof(double) * (CH_ROWS * CH_AS + CH_NB * CH_AS + CH_ROWS * CH_CS + CH_NB * CH_CS);

struct CholParams {
  int n, nr, ld, np, assemble;
  SnData S;
  const double* nuis;
  int n_nuis, ia, ib;
  double* W;
  size_t pt_stride;
  int* status;  // [np] != 0: not positive definite
};

__global__ void __launch_bounds__(256, 2) sn_chol_kernel(CholParams p) {
  extern __shared__ __align__(16) unsigned char ch_smem[];
  double* As = reinterpret_cast<double*>(ch_smem);  // [128][36]
  double* Bs = As + CH_ROWS * CH_AS;                 // [32][36]
  double* Cs = Bs + CH_NB * CH_AS;                   // [128][33]
  double* Ls = Cs + CH_ROWS * CH_CS;                 // [32][33]
  __shared__ int s_bad;
  const int pt = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n, nrows = p.n + p.nr, ld = p.ld;
  double* W = p.W + (size_t)pt * p.pt_stride;
  double alpha = 0, beta = 0;
  if (p.assemble) {
    alpha = p.ia >= 0 ? p.nuis[(size_t)pt * p.n_nuis + p.ia] : 0.0;
    beta = p.ib >= 0 ? p.nuis[(size_t)pt * p.n_nuis + p.ib] : 0.0;
  }
  const double cf[6] = {1.0, alpha * alpha, beta * beta, 2.0 * alpha, -2.0 * beta, -2.0 * alpha * beta};
  if (tid == 0) s_bad = 0;
  __syncthreads();

  for (int c0 = 0; c0 < n; c0 += CH_NB) {
    const int wcols = min(CH_NB, n - c0);
    for (int r0 = c0; r0 < nrows; r0 += CH_ROWS) {
      // ---- acc = sum_k L[r0+.., k] * L[c0+.., k], k < c0 (DMMA), warp = 16 rows x 32 cols
      double acc[2][4][2];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
      // The two L panels of chunk k0 + KC are loaded into registers while chunk k0 is multiplied (the global-load
      // latency used to sit exposed between the two barriers of every chunk: ncu long-scoreboard 12.8).
      constexpr int NA = CH_ROWS * (CH_KC / 2) / 256, NB2 = CH_NB * (CH_KC / 2) / 256;
      double2 pa[NA], pb[NB2];
      auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < NA; i++) {
          const int e = tid + i * 256;
          const int rr = e / (CH_KC / 2), kk = (e % (CH_KC / 2)) * 2;
          const int row = r0 + rr;
          pa[i] = (row < nrows) ? *reinterpret_cast<const double2*>(W + (size_t)row * ld + k0 + kk) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int i = 0; i < NB2; i++) {
          const int e = tid + i * 256;
          const int rr = e / (CH_KC / 2), kk = (e % (CH_KC / 2)) * 2;
          const int row = c0 + rr;
          pb[i] = (row < n) ? *reinterpret_cast<const double2*>(W + (size_t)row * ld + k0 + kk) : make_double2(0.0, 0.0);
        }
      };
      if (c0 > 0) fetch(0);
      for (int k0 = 0; k0 < c0; k0 += CH_KC) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NA; i++) {
          const int e = tid + i * 256;
          const int rr = e / (CH_KC / 2), kk = (e % (CH_KC / 2)) * 2;
          As[rr * CH_AS + kk] = pa[i].x;
          As[rr * CH_AS + kk + 1] = pa[i].y;
        }
#pragma unroll
        for (int i = 0; i < NB2; i++) {
          const int e = tid + i * 256;
          const int rr = e / (CH_KC / 2), kk = (e % (CH_KC / 2)) * 2;
          Bs[rr * CH_AS + kk] = pb[i].x;
          Bs[rr * CH_AS + kk + 1] = pb[i].y;
        }
        __syncthreads();
        if (k0 + CH_KC < c0) fetch(k0 + CH_KC);
#pragma unroll
        for (int kk = 0; kk < CH_KC / 4; kk++) {
          double a[2], b[4];
#pragma unroll
          for (int mi = 0; mi < 2; mi++) a[mi] = As[(warp * 16 + mi * 8 + (lane >> 2)) * CH_AS + kk * 4 + (lane & 3)];
#pragma unroll
          for (int ni = 0; ni < 4; ni++) b[ni] = Bs[(ni * 8 + (lane >> 2)) * CH_AS + kk * 4 + (lane & 3)];
#pragma unroll
          for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
      }
      // ---- Cs = A - acc  (A assembled from the constant blocks, or read from W)
#pragma unroll
      for (int mi = 0; mi < 2; mi++) {
        const int rr = warp * 16 + mi * 8 + (lane >> 2);
        const int row = r0 + rr;
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int cc = ni * 8 + (lane & 3) * 2 + e;
            const int col = c0 + cc;
            double av = 0.0;
            if (row < nrows && cc < wcols) {
              if (row < n && p.assemble) {
                const size_t o = (size_t)row * n + col;
#pragma unroll
                for (int m = 0; m < 6; m++)
                  if (p.S.cov[m]) av += cf[m] * p.S.cov[m][o];
                if (row == col) av += sn_diag(p.S, row, alpha, beta);
              } else {
                av = W[(size_t)row * ld + col];
              }
            }
            Cs[rr * CH_CS + cc] = av - acc[mi][ni][e];
          }
      }
      __syncthreads();
      if (r0 == c0) {
        // ---- factor the diagonal block (warp 0, lane = row)
        if (warp == 0) {
          double rowv[CH_NB];
#pragma unroll
          for (int j = 0; j < CH_NB; j++) rowv[j] = (lane < wcols && j < wcols) ? Cs[lane * CH_CS + j] : (lane == j ? 1.0 : 0.0);
          bool bad = false;
#pragma unroll
          for (int k = 0; k < CH_NB; k++) {
            double dkk = __shfl_sync(0xffffffffu, rowv[k], k);
            if (!(dkk > 0.0)) { bad = true; dkk = 1.0; }
            const double d = sqrt(dkk);
            const double lik = (lane >= k) ? ((lane == k) ? d : rowv[k] / d) : 0.0;
            rowv[k] = lik;
#pragma unroll
            for (int j = k + 1; j < CH_NB; j++) {
              const double ljk = __shfl_sync(0xffffffffu, lik, j);
              if (lane >= j) rowv[j] -= lik * ljk;
            }
          }
          if (bad && lane == 0) s_bad = 1;
#pragma unroll
          for (int j = 0; j < CH_NB; j++) Ls[lane * CH_CS + j] = (j <= lane) ? rowv[j] : 0.0;
          Ls[lane * CH_CS + CH_NB] = 1.0 / rowv[lane];  // rowv[lane] = L_lane,lane (1 beyond wcols)
        }
        __syncthreads();
      }
      // ---- rows of the diagonal block take L_jj; the others solve X L_jj^T = Cs row by row
      if (tid < CH_ROWS) {
        const int row = r0 + tid;
        if (row < nrows) {
          double* out = W + (size_t)row * ld + c0;
          if (row < c0 + wcols && r0 == c0) {
            for (int c = 0; c < wcols; c++) out[c] = Ls[tid * CH_CS + c];
          } else {
            // forward substitution, right-looking: once x_k is known every later column is updated independently,
            // so the dependent chain is 32 multiply-add steps instead of 32 divisions with growing dot products
            // (1 / L_kk is taken once per column; values, not indices)
            double x[CH_NB];
#pragma unroll
            for (int c = 0; c < CH_NB; c++) x[c] = Cs[tid * CH_CS + c];
#pragma unroll
            for (int k = 0; k < CH_NB; k++) {
              const double xk = x[k] * Ls[k * CH_CS + CH_NB];   // column CH_NB of Ls holds 1 / L_kk
              x[k] = xk;
#pragma unroll
              for (int c = k + 1; c < CH_NB; c++) x[c] -= xk * Ls[c * CH_CS + k];
            }
#pragma unroll
            for (int c = 0; c < CH_NB; c++)
              if (c < wcols) out[c] = x[c];
          }
        }
      }
      __syncthreads();
    }
    __threadfence_block();
  }
  if (tid == 0 && p.status) p.status[pt] = s_bad;
}

// ---- blocked Cholesky, second generation (the per-point path; V pre-assembled in W) -----------------------------------
// ncu of the kernel above (profiles/r01_sn_chol_ncu_full.txt): 0.23 of the DMMA peak, 7.9 barrier stalls per issued
// instruction - two __syncthreads per 32-wide k-chunk around a shared-memory copy of BOTH operands, 128-row passes so the
// column panel is re-staged six times for the early block columns, half of the threads idle in the panel solve.  Here:
//   * a warp owns 32 rows of the pass and takes its A fragments straight from global memory, two k-values per lane with
//     one 16-byte load (lane (r, c) holds k = 2c, 2c+1: the even and the odd values feed two DMMAs whose k-order is
//     permuted identically for A and B, which a sum over k does not see) - no shared-memory round trip, no barrier;
//   * only the 32 x K column panel L[c0.., 0..c0) goes through shared memory, 128 k-values per stage, double-buffered with
//     cp.async: ONE barrier per 128 k-values instead of two per 32;
//   * passes of 256 rows (8 warps x 32): the panel is staged at most three times, all 256 threads solve a row each;
//   * warps whose 32 rows lie past the matrix skip the k-loop (they only meet the barriers).
//   * NW warps per CTA (template): with 4 warps and 64 k-values per stage four CTAs share an SM, so the serial phases of one
//     matrix (diagonal-block factor, row solves, stage barriers) are covered by the k-loops of three others.
#ifndef CB200_C2_PF
#define CB200_C2_PF 1      // prefetch of the A rows: 0 off, 1 into L1, 2 into L2 (1 024 points: 9.33 -> 9.05 us/point; distance 64: 9.20, L2 / 128: 9.88)
#endif
#ifndef CB200_C2_PD
#define CB200_C2_PD 32     // prefetch distance in k-values
#endif
constexpr int C2_NB = 32, C2_CS = C2_NB + 1, C2_LS = C2_NB + 2;
template <int NW> struct C2Cfg {
  static constexpr int ROWS = 32 * NW;                    // rows per pass
  static constexpr int KS = NW >= 8 ? 128 : 64;           // k-values per stage of the column panel
  static constexpr int BST = KS + 8;
  static constexpr size_t BUF = sizeof(double) * 2 * C2_NB * BST;                     // two stages of the column panel
  static constexpr size_t CSB = sizeof(double) * ROWS * C2_CS;
  static constexpr size_t SCR = sizeof(double) * (NW - 1) * 1024;                     // k-split partial sums
  static constexpr size_t A0 = BUF > CSB ? (BUF > SCR ? BUF : SCR) : (CSB > SCR ? CSB : SCR);
  static constexpr size_t SMEM = A0 + sizeof(double) * C2_NB * C2_LS + sizeof(double) * 2 * C2_NB;  // + L_jj + column broadcast
};

__device__ __forceinline__ void c2_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void c2_cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int NW>
__global__ void __launch_bounds__(32 * NW, 16 / NW) sn_chol2_kernel(CholParams p) {
  constexpr int C2_ROWS = C2Cfg<NW>::ROWS, C2_KS = C2Cfg<NW>::KS, C2_BST = C2Cfg<NW>::BST, NT = 32 * NW;
  extern __shared__ __align__(16) unsigned char c2_smem[];
  double* Bs = reinterpret_cast<double*>(c2_smem);             // [2][32][C2_BST]; aliased by Cs [256][33] in the epilogue
  double* Cs = Bs;
  double* Ls = reinterpret_cast<double*>(c2_smem + C2Cfg<NW>::A0);
  __shared__ int s_bad;
  const int pt = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;   // fragment row / k-pair (A), n / k-pair (B), row / column pair (C)
  const int n = p.n, nrows = p.n + p.nr, ld = p.ld;
  double* W = p.W + (size_t)pt * p.pt_stride;
  if (tid == 0) s_bad = 0;
  __syncthreads();

  for (int c0 = 0; c0 < n; c0 += C2_NB) {
    const int wcols = min(C2_NB, n - c0);
    const int K = c0;
    const int tiles = (nrows - c0 + 31) >> 5, passes = (tiles + NW - 1) / NW;
    const int nsc = (K + C2_KS - 1) / C2_KS;
    for (int pass = 0; pass < passes; pass++) {
      // a pass with few row tiles (the late block columns, the tail pass of the early ones) splits the k-range of every
      // stage over SF warps per tile instead of leaving warps idle at the stage barriers; partial sums meet in shared memory
      const int tp = min(NW, tiles - pass * NW);
      int SF = 1;                            // largest power of two with SF * tp <= NW
      while (2 * SF * tp <= NW) SF *= 2;
      const int TPW = NW / SF, tslot = warp % TPW, kpart = warp / TPW;
      const int tile = pass * NW + tslot;
      const bool active = tslot < tp;
      const int row0 = c0 + tile * 32;
      double acc[4][4][2];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
      // stage sc of the column panel: rows c0 .. c0+31, k in [sc*128, sc*128 + kw)
      auto fill = [&](int sc) {
        const int kw = min(C2_KS, K - sc * C2_KS);
        double* dst = Bs + (size_t)(sc & 1) * C2_NB * C2_BST;
        const int per_row = kw >> 1;   // 16-byte pieces per row
        for (int e = tid; e < C2_NB * per_row; e += NT) {
          const int rr = e / per_row, kk = (e - rr * per_row) * 2;
          const int row = c0 + rr;
          if (row < nrows) c2_cp_async16(dst + rr * C2_BST + kk, W + (size_t)row * ld + sc * C2_KS + kk);
          else *reinterpret_cast<double2*>(dst + rr * C2_BST + kk) = make_double2(0.0, 0.0);
        }
      };
      if (nsc > 0) fill(0);
      for (int sc = 0; sc < nsc; sc++) {
        c2_cp_async_wait_all();
        __syncthreads();                       // stage sc landed; everybody is done with stage sc - 1
        if (sc + 1 < nsc) fill(sc + 1);
        if (active) {
          const int kw = min(C2_KS, K - sc * C2_KS);
          const double* bs = Bs + (size_t)(sc & 1) * C2_NB * C2_BST + fr * C2_BST + 2 * fc;
          const double* ag = W + (size_t)(row0 + fr) * ld + sc * C2_KS + 2 * fc;
#pragma unroll 2
          for (int k8 = 8 * kpart; k8 < kw; k8 += 8 * SF) {
            double2 a2[4], b2[4];
#if CB200_C2_PF
            // the A fragments come straight from global memory (23 MB per point of left-looking re-reads, half of them
            // DRAM misses): each lane prefetches the line of ONE of the warp's 32 rows a few k-steps ahead
            if (row0 + fc * 8 + fr < nrows) {
              const double* pa = ag + (size_t)fc * 8 * ld - 2 * fc + k8 + CB200_C2_PD;
#if CB200_C2_PF == 1
              asm volatile("prefetch.global.L1 [%0];\n" ::"l"(pa));
#else
              asm volatile("prefetch.global.L2 [%0];\n" ::"l"(pa));
#endif
            }
#endif
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
              a2[mi] = (row0 + mi * 8 + fr < nrows) ? *reinterpret_cast<const double2*>(ag + (size_t)mi * 8 * ld + k8) : make_double2(0.0, 0.0);
#pragma unroll
            for (int ni = 0; ni < 4; ni++) b2[ni] = *reinterpret_cast<const double2*>(bs + ni * 8 * C2_BST + k8);
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
              for (int ni = 0; ni < 4; ni++) {
                dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a2[mi].x, b2[ni].x);
                dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a2[mi].y, b2[ni].y);
              }
          }
        }
      }
      __syncthreads();                         // the panel stages are dead: Cs takes their place
      if (SF > 1) {                            // (block-uniform) add the k-parts: scratch [part - 1][tile slot][value][lane]
        double* scr = Cs;
        if (active && kpart > 0) {
          double* d = scr + ((size_t)((kpart - 1) * TPW + tslot) * 32) * 32 + lane;
#pragma unroll
          for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
              d[((mi * 4 + ni) * 2) * 32] = acc[mi][ni][0];
              d[((mi * 4 + ni) * 2 + 1) * 32] = acc[mi][ni][1];
            }
        }
        __syncthreads();
        if (active && kpart == 0) {
          for (int q = 1; q < SF; q++) {
            const double* d = scr + ((size_t)((q - 1) * TPW + tslot) * 32) * 32 + lane;
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
              for (int ni = 0; ni < 4; ni++) {
                acc[mi][ni][0] += d[((mi * 4 + ni) * 2) * 32];
                acc[mi][ni][1] += d[((mi * 4 + ni) * 2 + 1) * 32];
              }
          }
        }
        __syncthreads();                       // scratch read before Cs (same memory) is written
      }
      // ---- Cs = A - acc for this tile's 32 rows
      if (active && kpart == 0) {
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
          const int rr = tslot * 32 + mi * 8 + fr;
          const int row = row0 + mi * 8 + fr;
#pragma unroll
          for (int ni = 0; ni < 4; ni++) {
            const int cc = ni * 8 + 2 * fc;
            double2 av = make_double2(0.0, 0.0);
            if (row < nrows) av = *reinterpret_cast<const double2*>(W + (size_t)row * ld + c0 + cc);
            Cs[rr * C2_CS + cc] = (cc < wcols) ? av.x - acc[mi][ni][0] : 0.0;
            Cs[rr * C2_CS + cc + 1] = (cc + 1 < wcols) ? av.y - acc[mi][ni][1] : 0.0;
          }
        }
      }
      __syncthreads();
      if (pass == 0) {
        // ---- factor the diagonal block (warp 0, lane = row): rows 0..31 of this pass
        if (warp == 0) {
          double rowv[C2_NB];
#pragma unroll
          for (int j = 0; j < C2_NB; j++) rowv[j] = (lane < wcols && j < wcols) ? Cs[lane * C2_CS + j] : (lane == j ? 1.0 : 0.0);
          bool bad = false;
          double rinv = 1.0;
          // Column k of the block (one value per lane) is broadcast through shared memory: one store, then every lane reads
          // the pivot and the 31 - k values below it with 16-byte loads.  (The first version fetched each l_jk with its own
          // pair of shuffles: 2 x 496 dependent shuffles per block put 27 % of the kernel's time on this ONE warp while the
          // other seven waited at the barrier behind it - ncu source page, profiles/r02_sn_chol2_ncu_full.txt.)
          double* colb = Ls + C2_NB * C2_LS;   // [2][32], double-buffered: one __syncwarp per pivot
#pragma unroll
          for (int k = 0; k < C2_NB; k++) {
            double* col = colb + (k & 1) * C2_NB;
            col[lane] = rowv[k];
            __syncwarp();
            double dkk = col[k];
            if (!(dkk > 0.0)) { bad = true; dkk = 1.0; }
            const double rk = rsqrt(dkk);   // one reciprocal square root per pivot instead of a square root and a division
            const double d = dkk * rk;
            const double lik = (lane >= k) ? ((lane == k) ? d : rowv[k] * rk) : 0.0;
            rinv = (lane == k) ? rk : rinv;
            rowv[k] = lik;
            const double t = lik * rk;      // l_ik l_jk = (l_ik / sqrt(d_kk)) a_jk
            // unpredicated (entries above the diagonal are never read back), two columns per 16-byte load
#pragma unroll
            for (int j2 = (k + 1) & ~1; j2 < C2_NB; j2 += 2) {
              const double2 cj = *reinterpret_cast<const double2*>(col + j2);
              if (j2 > k) rowv[j2] -= t * cj.x;
              rowv[j2 + 1] -= t * cj.y;
            }
          }
          if (bad && lane == 0) s_bad = 1;
#pragma unroll
          for (int j = 0; j < C2_NB; j++) Ls[lane * C2_LS + j] = (j <= lane) ? rowv[j] : 0.0;
          Ls[lane * C2_LS + C2_NB] = rinv;   // 1 / L_lane,lane (1 beyond wcols)
        }
        __syncthreads();
      }
      // ---- every thread one row: the diagonal block's rows take L_jj, the others solve X L_jj^T = Cs
      {
        const int row = c0 + pass * C2_ROWS + tid;
        if (row < nrows) {
          double* out = W + (size_t)row * ld + c0;
          if (pass == 0 && tid < wcols) {
            for (int c = 0; c < wcols; c++) out[c] = Ls[tid * C2_LS + c];
          } else {   // (in the last block column the rows right below the diagonal block are the right-hand sides)
            double x[C2_NB];
#pragma unroll
            for (int c = 0; c < C2_NB; c++) x[c] = Cs[tid * C2_CS + c];
#pragma unroll
            for (int k = 0; k < C2_NB; k++) {
              const double xk = x[k] * Ls[k * C2_LS + C2_NB];
              x[k] = xk;
#pragma unroll
              for (int c = k + 1; c < C2_NB; c++) x[c] -= xk * Ls[c * C2_LS + k];
            }
#pragma unroll
            for (int c = 0; c < C2_NB; c++)
              if (c < wcols) out[c] = x[c];
          }
        }
      }
      __syncthreads();                         // Cs is dead before the next pass refills the stages
    }
    __threadfence_block();
  }
  if (tid == 0 && p.status) p.status[pt] = s_bad;
}

// chi^2 from the solved right-hand-side rows y_d, y_1, y_2 (each [n], row stride ld)
__global__ void __launch_bounds__(256) sn_final_kernel(int np, int n, int twoscriptm, const double* __restrict__ Y,
                                                       size_t pt_stride, int ld, const int* __restrict__ bad,
                                                       double* __restrict__ out, int out_stride) {
  __shared__ double sh[8];
  const int pt = blockIdx.x;
  const double* yd = Y + (size_t)pt * pt_stride;
  const double* y1 = yd + ld;
  const double* y2 = yd + 2 * (size_t)ld;
  double A = 0, B = 0, C = 0, D = 0, E = 0, F = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = yd[i], a1 = y1[i];
    A += d * d; B += d * a1; E += a1 * a1;
    if (twoscriptm) {
      const double a2 = y2[i];
      C += d * a2; D += a1 * a2; F += a2 * a2;
    }
  }
  A = block_sum_256(A, sh); B = block_sum_256(B, sh); E = block_sum_256(E, sh);
  if (twoscriptm) { C = block_sum_256(C, sh); D = block_sum_256(D, sh); F = block_sum_256(F, sh); }
  if (threadIdx.x == 0) {
    const double inv_twopi = 1.0 / (2 * 3.14159265358979323846264338328);
    double chisq;
    if (twoscriptm) {
      const double G = F - D * D / E;
      chisq = (G > 0) ? A + log(E * inv_twopi) + log(G * inv_twopi) - C * C / G - B * B * F / (E * G) + 2.0 * B * C * D / (E * G)
                      : 2e30;
    } else {
      chisq = A + log(E * inv_twopi) - B * B / E;
    }
    double r = chisq / 2;
    if (bad && bad[pt]) r = 1e30;
    out[(size_t)pt * out_stride] = r;
  }
}

// covariance independent of (alpha, beta): V^-1 cached at set-up; T = D V^-1 (DMMA GEMM) then per point
//   A = T.d, B = sum(T), E = sum(V^-1) (constant)
__global__ void __launch_bounds__(256) sn_final_cached_kernel(int np, int n, const double* __restrict__ T,
                                                              const double* __restrict__ Dm, double E,
                                                              double* __restrict__ out, int out_stride) {
  __shared__ double sh[8];
  const int pt = blockIdx.x;
  double A = 0, B = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double t = T[(size_t)pt * n + i];
    A += t * Dm[(size_t)pt * n + i];
    B += t;
  }
  A = block_sum_256(A, sh);
  B = block_sum_256(B, sh);
  if (threadIdx.x == 0) {
    const double inv_twopi = 1.0 / (2 * 3.14159265358979323846264338328);
    out[(size_t)pt * out_stride] = (A + log(E * inv_twopi) - B * B / E) / 2;
  }
}

}  // namespace cb200
