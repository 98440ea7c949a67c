// cosmob200 — shared device/host helpers (product code, sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>

namespace cb200 {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw cb200::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +  \
                             __FILE__ + ":" + std::to_string(__LINE__));                           \
  } while (0)

#define CB_LAUNCH_CHECK() CB_CUDA(cudaGetLastError())

template <class T>
struct DevBuf {  // owning device buffer
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    CB_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    n = count;
  }
  void zero(cudaStream_t s) { if (p) CB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) {
    alloc(count);
    CB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& v, cudaStream_t s) { upload(v.data(), v.size(), s); }
};

constexpr double kPi = 3.14159265358979323846264338328;
constexpr double kTwoPi = 2 * kPi;
constexpr double kFourPi = 4 * kPi;

// up to 8 uniformly sampled stretches of a grid: {lo, hi, step, first(1-based)}; linear only
struct LinSegs {
  int n;
  double highest;
  int npoints;
  double seg[8][4];
  double inv_step[8];  // 1/step, for interpolation weights only (never for index arithmetic)
};

// 1-based index of the last sample <= v, with the reference's arithmetic (camb/utils.F90:81-111):
// truncation of a true IEEE division, no FMA.
__device__ __forceinline__ int lin_index_of(const LinSegs& g, double v) {
#pragma unroll 1
  for (int r = 0; r < g.n; r++) {
    double lo = g.seg[r][0];
    if (v < g.seg[r][1] && v >= lo) return (int)g.seg[r][3] + (int)(__ddiv_rn(__dsub_rn(v, lo), g.seg[r][2]));
  }
  if (v >= g.highest) return g.npoints;
  return 1;
}

// same lookup, also returning the two grid values bracketing v, rebuilt arithmetically exactly as the grid was
// materialised (lo + step*j without FMA; the last sample of a stretch is the next stretch's lo): no table load.
__device__ __forceinline__ int lin_locate(const LinSegs& g, double v, double& x0, double& x1, double& inv_h) {
#pragma unroll 1
  for (int r = 0; r < g.n; r++) {
    const double lo = g.seg[r][0], hi = g.seg[r][1];
    if (v < hi && v >= lo) {
      const double step = g.seg[r][2];
      const int first = (int)g.seg[r][3];
      const int j = (int)(__ddiv_rn(__dsub_rn(v, lo), step));
      const int nseg = ((r + 1 < g.n) ? (int)g.seg[r + 1][3] : g.npoints) - first;
      x0 = __dadd_rn(lo, __dmul_rn(step, (double)j));
      x1 = (j + 1 < nseg) ? __dadd_rn(lo, __dmul_rn(step, (double)(j + 1))) : hi;
      inv_h = g.inv_step[r];
      return first + j;
    }
  }
  x0 = x1 = g.highest;
  inv_h = 0;
  return g.npoints;
}

// lin_locate with the stretches tried from the last one down: the stretches are disjoint, so the result is the same;
// the Bessel abscissae of almost every (q, tau) pair lie in the last (coarsest) stretch, found on the first try
__device__ __forceinline__ int lin_locate_desc(const LinSegs& g, double v, double& x0, double& x1, double& inv_h) {
#pragma unroll 1
  for (int r = g.n - 1; r >= 0; r--) {
    const double lo = g.seg[r][0], hi = g.seg[r][1];
    if (v < hi && v >= lo) {
      const double step = g.seg[r][2];
      const int first = (int)g.seg[r][3];
      const int j = (int)(__ddiv_rn(__dsub_rn(v, lo), step));
      const int nseg = ((r + 1 < g.n) ? (int)g.seg[r + 1][3] : g.npoints) - first;
      x0 = __dadd_rn(lo, __dmul_rn(step, (double)j));
      x1 = (j + 1 < nseg) ? __dadd_rn(lo, __dmul_rn(step, (double)(j + 1))) : hi;
      inv_h = g.inv_step[r];
      return first + j;
    }
  }
  x0 = x1 = g.highest;
  inv_h = 0;
  return g.npoints;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cb200
