// Stand-alone probe: which shared-memory descriptor encoding does tcgen05.mma.kind::tf32 want for MN-major operands
// loaded by TMA as [32 K rows] x [32 channel] SWIZZLE_128B boxes?   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (long long i = 0; i < (1LL << 24); ++i) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
struct alignas(64) Maps { CUtensorMap a, b; };
struct Variant { uint32_t lbo, sbo, kadv, amaj, bmaj; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Maps maps, Variant v, float* out /*[128][64]*/) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t sa = smem_u32(smem), sb = sa + 4 * 4096;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full)), "r"(6u * 4096u) : "memory");
    for (int g = 0; g < 4; ++g)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sa + g * 4096), "l"(&maps.a), "r"(smem_u32(&full)), "r"(g * 32), "r"(0) : "memory");
    for (int g = 0; g < 2; ++g)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sb + g * 4096), "l"(&maps.b), "r"(smem_u32(&full)), "r"(g * 32), "r"(0) : "memory");
    mbar_wait(&full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (v.amaj << 15) | (v.bmaj << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = ((uint64_t)(((sa + k * v.kadv) >> 4) & 0x3FFF)) | ((uint64_t)((v.lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((v.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
      uint64_t db = ((uint64_t)(((sb + k * v.kadv) >> 4) & 0x3FFF)) | ((uint64_t)((v.lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((v.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
      uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
  }
  mbar_wait(&done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t x[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]), "=r"(x[9]),
                   "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[r * 64 + c0 + i] = __uint_as_float(x[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  const int K = 32, M = 128, N = 64;
  std::vector<float> A(K * M), B(K * N), ref(M * N, 0.f);
  for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) A[k * M + m] = (float)(((k * 7 + m * 3) % 13) - 6);
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) B[k * N + n] = (float)(((k * 5 + n) % 11) - 5);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[k * M + m] * B[k * N + n]; ref[m * N + n] = s; }
  float *dA, *dB, *dO; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  Maps maps;
  { cuuint64_t gd[2] = {(cuuint64_t)M, (cuuint64_t)K}; cuuint64_t gs[1] = {(cuuint64_t)M * 4}; cuuint32_t bx[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
    enc(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  { cuuint64_t gd[2] = {(cuuint64_t)N, (cuuint64_t)K}; cuuint64_t gs[1] = {(cuuint64_t)N * 4}; cuuint32_t bx[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
    enc(&maps.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  Variant vs[] = {{4096, 1024, 1024, 1, 1}, {1024, 4096, 1024, 1, 1}, {4096, 1024, 32, 1, 1}, {4096, 128, 1024, 1, 1}, {128, 4096, 1024, 1, 1},
                  {4096, 1024, 1024, 0, 0}, {0, 1024, 32, 0, 0}, {4096, 256, 1024, 1, 1}, {256, 4096, 1024, 1, 1}};
  std::vector<float> out(M * N);
  for (auto& v : vs) {
    cudaMemset(dO, 0xFF, M * N * 4);
    probe<<<1, 128, 64 * 1024>>>(maps, v, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant lbo=%u sbo=%u kadv=%u maj=%u%u : CUDA error %s\n", v.lbo, v.sbo, v.kadv, v.amaj, v.bmaj, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxabs = 0; int nz = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(out[i] - ref[i])); maxabs = fmax(maxabs, fabs(out[i])); nz += out[i] != 0.f; }
    printf("variant lbo=%4u sbo=%4u kadv=%4u maj=%u%u : max|err| = %.3f  max|out| = %.1f  nonzero = %d   out[0..3] = %.1f %.1f %.1f %.1f (ref %.1f %.1f %.1f %.1f)\n",
           v.lbo, v.sbo, v.kadv, v.amaj, v.bmaj, maxerr, maxabs, nz, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
