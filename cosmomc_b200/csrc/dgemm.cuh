// FP64 GEMM on the tensor pipe (DMMA, mma.sync.m8n8k4.f64) for the dense contractions of the path:
// lensing correlation sums, band-power binning, chi^2 quadratic forms.  tcgen05 has no FP64 kind, so the
// FP64 tensor work on sm_100a goes through the legacy mma.sync DMMA instruction.
//
//   C[M,N] = alpha * A[M,K] * op(B) (+ C if accumulate)     row-major, arbitrary leading dimensions
//   op(B) = B[K,N] (transB = false) or B[N,K]^T (transB = true)
#pragma once
#include "common.cuh"

namespace cb200 {

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_BK = 16;
constexpr int GEMM_AS = GEMM_BK + 4;  // smem strides chosen so the 16 lanes of a half-warp hit 16 banks pairs
constexpr int GEMM_BS = GEMM_BN + 4;

template <bool TRANSB, bool ACCUM>
__global__ void __launch_bounds__(128) dgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                                                    int lda, const double* __restrict__ B, int ldb,
                                                    double* __restrict__ C, int ldc) {
  __shared__ double As[GEMM_BM * GEMM_AS];
  __shared__ double Bs[GEMM_BK * GEMM_BS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // global->register staging: A tile 64x16 (thread: row tid/2, 8 cols), B tile 16x64 (thread: row tid/8, 8 cols)
  double ra[8], rb[8];
  const int a_row = tid >> 1, a_col = (tid & 1) * 8;
  const int b_row = tid >> 3, b_col = (tid & 7) * 8;

  auto load_tiles = [&](int k0) {
    const int gm = m0 + a_row;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int gk = k0 + a_col + i;
      ra[i] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.0;
    }
    const int gk = k0 + b_row;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int gn = n0 + b_col + i;
      if (TRANSB) rb[i] = (gk < K && gn < N) ? B[(size_t)gn * ldb + gk] : 0.0;
      else rb[i] = (gk < K && gn < N) ? B[(size_t)gk * ldb + gn] : 0.0;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 8; i++) As[a_row * GEMM_AS + a_col + i] = ra[i];
#pragma unroll
    for (int i = 0; i < 8; i++) Bs[b_row * GEMM_BS + b_col + i] = rb[i];
  };

  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += GEMM_BK) {
    __syncthreads();
    store_tiles();
    __syncthreads();
    if (k0 + GEMM_BK < K) load_tiles(k0 + GEMM_BK);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK / 4; kk++) {
      double a[4], b[4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) a[mi] = As[(wm * 32 + mi * 8 + (lane >> 2)) * GEMM_AS + kk * 4 + (lane & 3)];
#pragma unroll
      for (int ni = 0; ni < 4; ni++) b[ni] = Bs[(kk * 4 + (lane & 3)) * GEMM_BS + wn * 32 + ni * 8 + (lane >> 2)];
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
#pragma unroll
  for (int mi = 0; mi < 4; mi++) {
    int gm = m0 + wm * 32 + mi * 8 + (lane >> 2);
    if (gm >= M) continue;
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
      int gn = n0 + wn * 32 + ni * 8 + (lane & 3) * 2;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        if (gn + e < N) {
          double v = alpha * acc[mi][ni][e];
          size_t o = (size_t)gm * ldc + gn + e;
          C[o] = ACCUM ? C[o] + v : v;
        }
      }
    }
  }
}

inline void dgemm(cudaStream_t s, bool transB, bool accumulate, int M, int N, int K, double alpha, const double* A,
                  int lda, const double* B, int ldb, double* C, int ldc, long long* launches = nullptr) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((N + GEMM_BN - 1) / GEMM_BN, (M + GEMM_BM - 1) / GEMM_BM);
  if (transB) {
    if (accumulate) dgemm_kernel<true, true><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, C, ldc);
    else dgemm_kernel<true, false><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, C, ldc);
  } else {
    if (accumulate) dgemm_kernel<false, true><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, C, ldc);
    else dgemm_kernel<false, false><<<grid, 128, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, C, ldc);
  }
  CB_LAUNCH_CHECK();
  if (launches) ++*launches;
}

}  // namespace cb200
