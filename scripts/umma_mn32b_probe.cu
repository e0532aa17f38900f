// Stand-alone probe (round 2): tcgen05.mma.kind::tf32 with MN-MAJOR operands, i.e. operands stored [K rows][MN contiguous]
// -- exactly how the engine stores activations ([pixel rows][channels]) -- which is what a tensor-core weight gradient
// needs (reduction over pixels).  CUTLASS (sm100_common.inl) says: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout": TMA swizzle CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B + descriptor layout type 1
// (SWIZZLE_128B_BASE32B).  Round 1's probe (umma_mn_probe.cu) only tried layout type 2 and saw zeros.
// Questions answered here (each variant prints whether the result matches under several TMEM-mapping hypotheses):
//   H1  does MN-major TF32 work with (ATOM_32B, layout type 1)?  LBO / SBO semantics?
//   H2  may the descriptor start at an arbitrary ROW of the tile (start += row * 128 B, base_offset = 0)?  (halo trick)
//   H3  where does an M = 64 accumulator live in TMEM (lanes / columns)?
//   H4  N = 128 with the B operand = [B_hi tiles ; B_lo tiles] (4 MN atoms at a uniform LBO stride)
// nvcc -gencode arch=compute_100a,code=sm_100a -o umma_mn32b_probe umma_mn32b_probe.cu -lcuda
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
struct Variant {
  uint32_t layout;      // descriptor layout type (1 = 128B_BASE32B, 2 = 128B)
  uint32_t lbo, sbo;    // bytes
  uint32_t kadv;        // bytes added to the start address per K = 8 step
  uint32_t amaj, bmaj;  // 1 = MN-major
  uint32_t M, N;        // instruction shape
  uint32_t ksteps;      // number of K = 8 steps
  uint32_t row_off;     // A descriptor starts this many rows into the tile
  uint32_t b_atoms;     // B tiles loaded (32-wide column blocks): 2 for N = 64, 4 for N = 128
};
#define KR 48          // rows per tile (K extent available)
#define TILE (KR * 128)

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Maps maps, Variant v, float* out /*[128 lanes][128 cols]*/) {
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  // A tiles: 4 x TILE (channel blocks 0..3 of a [KR][128] matrix); B tiles: 4 x TILE
  const uint32_t sa = smem_u32(smem), sb = sa + 4 * TILE;
  // zero TMEM first (so that untouched lanes / columns read as 0): every warp stores zeros to its lane quarter
  {
    for (int c0 = 0; c0 < 128; c0 += 8) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + c0), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full)), "r"((4u + v.b_atoms) * TILE) : "memory");
    for (int g = 0; g < 4; ++g)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sa + g * TILE), "l"(&maps.a), "r"(smem_u32(&full)), "r"(g * 32), "r"(0) : "memory");
    for (int g = 0; g < (int)v.b_atoms; ++g)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sb + g * TILE), "l"(&maps.b), "r"(smem_u32(&full)), "r"(g * 32), "r"(0) : "memory");
    mbar_wait(&full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (v.amaj << 15) | (v.bmaj << 16) | ((v.N >> 3) << 17) | ((v.M >> 4) << 24);
    for (uint32_t k = 0; k < v.ksteps; ++k) {
      const uint32_t a0 = sa + v.row_off * 128u + k * v.kadv, b0 = sb + k * v.kadv;
      uint64_t da = ((uint64_t)((a0 >> 4) & 0x3FFF)) | ((uint64_t)((v.lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((v.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)v.layout << 61);
      uint64_t db = ((uint64_t)((b0 >> 4) & 0x3FFF)) | ((uint64_t)((v.lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((v.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)v.layout << 61);
      uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
  }
  mbar_wait(&done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t x[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]), "=r"(x[9]),
                   "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[r * 128 + c0 + i] = __uint_as_float(x[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeFn enc, CUtensorMap* m, float* base, CUtensorMapSwizzle sw) {
  cuuint64_t gd[2] = {128, KR}; cuuint64_t gs[1] = {128 * 4}; cuuint32_t bx[2] = {32, KR}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
}

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  // A[j][m], B[j][n]: [KR][128].  Data set 0 ("unique"): D[m][n] = 128 (m + 1) + (n + 1) from rows row_off, row_off + 1
  // of A against rows 0, 1 of B.  Data set 1: small pseudo-random integers (exact in tf32).
  float *dA, *dB, *dO; cudaMalloc(&dA, KR * 128 * 4); cudaMalloc(&dB, KR * 128 * 4); cudaMalloc(&dO, 128 * 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * TILE + 2048);
  struct Case { const char* name; int data; int tma_sw; Variant v; };
  const uint32_t T = TILE;
  Case cases[] = {
    {"H1 type1 ATOM32B lbo=T sbo=512 M64 N64 unique", 0, 1, {1, T, 512, 1024, 1, 1, 64, 64, 1, 0, 2}},
    {"H1b type1 ATOM32B lbo=512 sbo=T (swapped)", 0, 1, {1, 512, T, 1024, 1, 1, 64, 64, 1, 0, 2}},
    {"H1c type2 SW128 lbo=T sbo=1024 (round-1 layout)", 0, 0, {2, T, 1024, 1024, 1, 1, 64, 64, 1, 0, 2}},
    {"H1d type1 SW128-TMA (mismatch control)", 0, 0, {1, T, 512, 1024, 1, 1, 64, 64, 1, 0, 2}},
    {"H3 type1 M128 N64 unique", 0, 1, {1, T, 512, 1024, 1, 1, 128, 64, 1, 0, 2}},
    {"K2 type1 M64 N64 random K=16", 1, 1, {1, T, 512, 1024, 1, 1, 64, 64, 2, 0, 2}},
    {"K4 type1 M64 N64 random K=32", 1, 1, {1, T, 512, 1024, 1, 1, 64, 64, 4, 0, 2}},
    {"H2 row_off=1 random K=16", 1, 1, {1, T, 512, 1024, 1, 1, 64, 64, 2, 1, 2}},
    {"H2 row_off=2 random K=16", 1, 1, {1, T, 512, 1024, 1, 1, 64, 64, 2, 2, 2}},
    {"H2 row_off=3 random K=16", 1, 1, {1, T, 512, 1024, 1, 1, 64, 64, 2, 3, 2}},
    {"H2 row_off=5 random K=32", 1, 1, {1, T, 512, 1024, 1, 1, 64, 64, 4, 5, 2}},
    {"H2 row_off=15 random K=32", 1, 1, {1, T, 512, 1024, 1, 1, 64, 64, 4, 15, 2}},
    {"H4 N=128 (4 B atoms) M64 random K=16", 1, 1, {1, T, 512, 1024, 1, 1, 64, 128, 2, 0, 4}},
    {"H4b N=128 M128 random K=16 row_off=6", 1, 1, {1, T, 512, 1024, 1, 1, 128, 128, 2, 6, 4}},
    {"H5 N=48 M64 random K=16 (F=48 layers)", 1, 1, {1, T, 512, 1024, 1, 1, 64, 48, 2, 0, 2}},
    {"H6 N=96 M64 random K=16 (stacked F=48: needs atoms 0,1,2,3 = hi0 hi1 lo0 lo1 -> cols 0-47, 64-111?)", 1, 1, {1, T, 512, 1024, 1, 1, 64, 96, 2, 0, 4}},
  };
  std::vector<float> A(KR * 128), B(KR * 128), out(128 * 128);
  for (auto& cs : cases) {
    const Variant& v = cs.v;
    for (auto& x : A) x = 0.f;
    for (auto& x : B) x = 0.f;
    if (cs.data == 0) {
      for (int m = 0; m < 128; ++m) { A[(v.row_off + 0) * 128 + m] = (float)(m + 1); A[(v.row_off + 1) * 128 + m] = 1.f; }
      for (int n = 0; n < 128; ++n) { B[0 * 128 + n] = 128.f; B[1 * 128 + n] = (float)(n + 1); }
    } else {
      for (int j = 0; j < KR; ++j) for (int m = 0; m < 128; ++m) A[j * 128 + m] = (float)(((j * 7 + m * 3 + (j * m) % 5) % 13) - 6);
      for (int j = 0; j < KR; ++j) for (int n = 0; n < 128; ++n) B[j * 128 + n] = (float)(((j * 5 + n + (j * n) % 3) % 11) - 5);
    }
    const int K = 8 * v.ksteps;
    std::vector<float> ref(128 * 128, 0.f);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[(v.row_off + k) * 128 + m] * B[k * 128 + n]; ref[m * 128 + n] = s; }
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    Maps maps;
    make_map(enc, &maps.a, dA, cs.tma_sw ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
    make_map(enc, &maps.b, dB, cs.tma_sw ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
    cudaMemset(dO, 0, 128 * 128 * 4);
    probe<<<1, 128, 8 * TILE + 2048>>>(maps, v, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-60s : CUDA error %s\n", cs.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dO, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    // mapping hypotheses (m, n) -> (lane, col)
    const int M = v.M, N = v.N;
    auto check = [&](const char* nm, auto f) {
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        int lane, col; f(m, n, lane, col);
        if (lane < 0 || lane >= 128 || col < 0 || col >= 128) { bad++; continue; }
        const double d = fabs(out[lane * 128 + col] - ref[m * 128 + n]);
        maxerr = fmax(maxerr, d); bad += d > 0.5;
      }
      printf("    %-44s max|err| %.1f  mismatches %d / %d\n", nm, maxerr, bad, M * N);
      return bad == 0;
    };
    int nz = 0; for (auto x : out) nz += x != 0.f;
    printf("%s  [nonzero TMEM cells: %d]\n", cs.name, nz);
    bool ok = false;
    ok |= check("lane=m col=n", [&](int m, int n, int& l, int& c) { l = m; c = n; });
    if (M == 64) {
      ok |= check("lane=32*(m/16)+m%16 col=n", [&](int m, int n, int& l, int& c) { l = 32 * (m / 16) + m % 16; c = n; });
      ok |= check("lane=32*(m/16)+m%16+16*(n>=N/2) col=n%(N/2)", [&](int m, int n, int& l, int& c) { l = 32 * (m / 16) + m % 16 + 16 * (n >= N / 2); c = n % (N / 2); });
      ok |= check("lane=m+64*(n>=N/2) col=n%(N/2)", [&](int m, int n, int& l, int& c) { l = m + 64 * (n >= N / 2); c = n % (N / 2); });
      ok |= check("lane=32*(m/16)+m%16+16*(n%2) col=n/2", [&](int m, int n, int& l, int& c) { l = 32 * (m / 16) + m % 16 + 16 * (n % 2); c = n / 2; });
    }
    if (!ok && cs.data == 0) {
      // decode where things landed: value = 128 (m + 1) + (n + 1)
      printf("    decoded (lane, col) -> (m, n) samples:");
      int shown = 0;
      for (int l = 0; l < 128 && shown < 24; ++l) for (int c = 0; c < 128 && shown < 24; ++c) {
        const float x = out[l * 128 + c];
        if (x != 0.f && (l % 16 == 0 || l % 16 == 1 || l % 16 == 15) && (c < 2 || c == 32 || c == 63)) {
          const int iv = (int)lrintf(x); printf(" (%d,%d)->(%d,%d)", l, c, iv / 128 - 1, iv % 128 - 1); shown++;
        }
      }
      printf("\n");
    }
  }
  return 0;
}
