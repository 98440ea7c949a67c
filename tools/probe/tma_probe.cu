// Dev probe (GPU): 2-D tensor-map TMA of FP64 boxes, the configuration project4_kernel uses for the raw source rows.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const CUtensorMap* tm, int c0, int c1, int bytes, double* out, int mode, const __grid_constant__ CUtensorMap tmp) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + 4096);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(su32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(su32(bar)), "r"(bytes) : "memory");
    const CUtensorMap* t = mode ? &tmp : tm;
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(su32(sm)),
                 "l"(t), "r"(c0), "r"(c1), "r"(su32(bar)) : "memory");
  }
  asm volatile("{\n .reg .pred p;\nW:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D;\n bra W;\nD:\n}\n" ::"r"(su32(bar)) : "memory");
  for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) out[i] = reinterpret_cast<double*>(sm)[i];
}
int main(int argc, char** argv) {
  const int NK = argc > 1 ? atoi(argv[1]) : 224, bx = argc > 2 ? atoi(argv[2]) : 8, by = argc > 3 ? atoi(argv[3]) : 12;
  const int mode = argc > 4 ? atoi(argv[4]) : 0, c0 = argc > 5 ? atoi(argv[5]) : 3, c1 = argc > 6 ? atoi(argv[6]) : 30;
  const size_t rows = 4096;
  std::vector<double> h(rows * NK);
  for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
  double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 8192);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)NK, rows}, str[1] = {(cuuint64_t)NK * 8};
  cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by}, es[2] = {1, 1};
  CUresult r = ((enc_t)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d  NK=%d box=%dx%d mode=%d c=(%d,%d)\n", (int)r, NK, bx, by, mode, c0, c1);
  CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  probe<<<1, 32, 8192>>>(dtm, c0, c1, bx * by * 8, o, mode, tm);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> ho(bx * by); cudaMemcpy(ho.data(), o, bx * by * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < by; y++) for (int x = 0; x < bx; x++) {
      double want = (c0 + x < NK) ? (double)((size_t)(c1 + y) * NK + c0 + x) : 0.0;
      if (ho[y * bx + x] != want) bad++;
    }
    printf("mismatches: %d (first row: %.0f %.0f %.0f)\n", bad, ho[0], ho[1], ho[2]);
  }
  return 0;
}
