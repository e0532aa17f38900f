// tcgen05 / TMA implicit-GEMM 3x3 convolution for sm_100a, fp32-faithful via the 3xTF32 operand split.
//
//   out[j, n] = sum_src sum_tap sum_k A_src[j +/- s_tap, k] * B_src[tap][n][k]        (fp32 result)
//
// Used for the forward conv, the tangent conv (two operand pairs), dgrad and tangent dgrad of blocks
// l >= 1 (reference meta_neural_network_architectures.py:89-97 and its autograd derivatives).
//
// Why it maps to plain 2-D TMA tiles: activations live on the zero-padded pixel grid (common.cuh), so the A
// operand of filter tap (ky,kx) is the SAME [rows, C] matrix shifted by s_tap rows -- every (tap, k-chunk)
// stage is one 128 x 32 fp32 box (SWIZZLE_128B) per operand half, out-of-range rows are zero-filled by TMA.
//
// Precision: single-pass TF32 is not acceptable for this path (40-130 % meta-gradient
// error).  Every operand x is pre-split by its producer kernel into hi = rna_tf32(x), lo = rna_tf32(x - hi);
// the kernel accumulates A_hi*B_hi and (A_lo*B_hi + A_hi*B_lo) in SEPARATE fp32 TMEM accumulators.
// The tensor core's fp32 accumulation truncates when it aligns addends, so error grows with the number of
// sequential accumulations into one accumulator (measured: one accumulator for all 216 MMAs of a 64-channel
// layer gave ~5x the fp32-FFMA error).  The big term is therefore spread round-robin over 4 accumulators
// (18 accumulations each instead of 216), the small terms get a fifth, and the epilogue adds the five with
// IEEE fp32 adds.
//
// CTA = one 128-row M tile x all N (<= 64) columns.  Warp roles: warp 0 = TMA producer (one thread),
// warp 1 = TMEM allocator + MMA issuer (one thread), warps 2..5 = epilogue (TMEM -> registers -> shared ->
// coalesced global store, + bias, + fp64 BatchNorm statistics).  4-stage smem ring, mbarrier full/empty,
// tcgen05.commit frees stages and publishes the accumulator.
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = 128 * 128;          // 128 rows x 32 fp32 (one 128B swizzle span per row)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a mis-programmed pipeline traps (error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (long long spin = 0; spin < (1LL << 26); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused: one swizzle atom along K) | [32,46) SBO >> 4 = 1024 B
//   (8 rows x 128 B) | [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ int tap_shift_tc(int tap, int gw) { return (tap / 3 - 1) * gw + (tap % 3 - 1); }

template <int NCOLS>
__global__ void __launch_bounds__(192, 1) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcConvArgs a) {
  constexpr int B_BYTES = NCOLS * 128;
  constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
  constexpr int NACC = 5;                       // 4 x hi*hi (round-robin over k-steps) + 1 x (lo*hi + hi*lo)
  constexpr int TMEM_COLS = NCOLS <= 16 ? 128 : (NCOLS <= 48 ? 256 : 512);   // power of two >= NACC * NCOLS
  constexpr int PITCH = NCOLS + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int row_ok[128];
  __shared__ double sred[2][NCOLS][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int task = blockIdx.y;
  const int j0 = blockIdx.x * 128;
  const int kchunks = (a.kc + 31) >> 5;        // a ragged last chunk (kc = 16 or 48) is zero-filled by TMA
  const int it0 = 9 * kchunks;
  const int nit = a.nsrc * it0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int it = 0; it < nit; ++it) {
        const int stage = it % TC_STAGES;
        const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        const int s = it / it0;
        const int local = it - s * it0;
        const int tap = local / kchunks;
        const int kc0 = (local - tap * kchunks) << 5;
        const int arow = a.a_row_base[s] + task * a.a_task_rows[s] + j0 + a.sign[s] * tap_shift_tc(tap, a.gw);
        const int brow = a.b_row_base[s] + task * a.b_task_rows[s] + tap * NCOLS;
        const uint32_t st = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
        tma_load_2d(st, &maps.m[s * 4 + 0], &full_bar[stage], kc0, arow);
        tma_load_2d(st + TC_A_BYTES, &maps.m[s * 4 + 1], &full_bar[stage], kc0, arow);
        tma_load_2d(st + 2 * TC_A_BYTES, &maps.m[s * 4 + 2], &full_bar[stage], kc0, brow);
        tma_load_2d(st + 2 * TC_A_BYTES + B_BYTES, &maps.m[s * 4 + 3], &full_bar[stage], kc0, brow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major,
      // N>>3 at bit 17, M>>4 at bit 24
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int it = 0; it < nit; ++it) {
        const int stage = it % TC_STAGES;
        const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)stage * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) {          // 4 x UMMA_K(8 tf32 = 32 B) per 128 B swizzle span
          const uint64_t ah = make_desc_sw128(st + k * 32);
          const uint64_t al = make_desc_sw128(st + TC_A_BYTES + k * 32);
          const uint64_t bh = make_desc_sw128(st + 2 * TC_A_BYTES + k * 32);
          const uint64_t bl = make_desc_sw128(st + 2 * TC_A_BYTES + B_BYTES + k * 32);
          const int kstep = it * 4 + k;
          tc_mma_tf32(tmem_base + 4 * NCOLS, al, bh, idesc, kstep > 0 ? 1u : 0u);
          tc_mma_tf32(tmem_base + 4 * NCOLS, ah, bl, idesc, 1u);
          tc_mma_tf32(tmem_base + (uint32_t)(kstep & 3) * NCOLS, ah, bh, idesc, kstep >= 4 ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);            // frees this smem stage once the MMAs above have read it
      }
      tc_commit(&accum_bar);                     // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int et = threadIdx.x - 64;             // 0..127
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // tile row owned by this thread in TMEM
    {
      const int row = j0 + et;
      int ok = 0;
      if (row < a.rows) {
        const int rr = row % a.G;
        const int yy = rr / a.gw, xx = rr - yy * a.gw;
        ok = (yy >= 1 && yy <= a.h && xx >= 1 && xx <= a.w) ? 1 : 0;
      }
      row_ok[et] = ok;
    }
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    float* tile = reinterpret_cast<float*>(smem);  // all MMAs have completed: the stage ring is free
    const float* bias = a.bias ? a.bias + (long long)task * a.bias_stride : nullptr;
#pragma unroll
    for (int c0 = 0; c0 < NCOLS; c0 += 16) {
      uint32_t v0[16], v1[16], v2[16], v3[16], v4[16];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      tc_ld16(ta, v0);
      tc_ld16(ta + NCOLS, v1);
      tc_ld16(ta + 2 * NCOLS, v2);
      tc_ld16(ta + 3 * NCOLS, v3);
      tc_ld16(ta + 4 * NCOLS, v4);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float big = (__uint_as_float(v0[i]) + __uint_as_float(v1[i])) + (__uint_as_float(v2[i]) + __uint_as_float(v3[i]));
        tile[r * PITCH + c0 + i] = (big + __uint_as_float(v4[i])) + (bias ? bias[c0 + i] : 0.f);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float* out = a.out + (long long)task * a.out_stride;
    const float* zh = a.zh ? a.zh + (long long)task * a.zh_stride : nullptr;
    constexpr int PARTS = 128 / NCOLS;           // 2 for 64 and 48, 4 for 32, 8 for 16
    const int col = et % NCOLS, part = et / NCOLS;
    double s1 = 0.0, s2 = 0.0;
    for (int idx = et; idx < 128 * NCOLS; idx += 128) {
      const int rr = idx / NCOLS, cc = idx - rr * NCOLS;
      const int row = j0 + rr;
      if (row < a.rows) out[(long long)row * NCOLS + cc] = tile[rr * PITCH + cc];
    }
    if (a.mode != CONV_PLAIN && part < PARTS) {
      for (int rr = part; rr < 128; rr += PARTS) {
        if (row_ok[rr]) {
          const float v = tile[rr * PITCH + col];
          if (a.mode == CONV_FWD_STATS) { s1 += (double)v; s2 += (double)v * (double)v; }
          else {
            const float zv = zh[(long long)(j0 + rr) * NCOLS + col];
            s1 += (double)v; s2 += (double)zv * (double)v;
          }
        }
      }
      if (part < 2) { sred[part][col][0] = s1; sred[part][col][1] = s2; }
    }
    if (a.mode != CONV_PLAIN) {
      if (PARTS > 2) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (part >= 2 && part < PARTS) { atomicAdd(&sred[part & 1][col][0], s1); atomicAdd(&sred[part & 1][col][1], s2); }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < NCOLS * 2) {
        const int c = et >> 1, which = et & 1;
        double* stats = a.stats + (long long)task * a.stats_stride;
        atomicAdd(&stats[c * 2 + which], sred[0][c][which] + sred[1][c][which]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace

size_t tc_conv_smem_bytes(int ncols) { return (size_t)TC_STAGES * (2 * TC_A_BYTES + 2 * ncols * 128) + 1024; }

int tc_conv_prepare() {
  cudaError_t e1 = cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_conv_smem_bytes(64));
  cudaError_t e2 = cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_conv_smem_bytes(32));
  cudaError_t e3 = cudaFuncSetAttribute(conv_tc_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_conv_smem_bytes(48));
  cudaError_t e4 = cudaFuncSetAttribute(conv_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_conv_smem_bytes(16));
  return (e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess && e4 == cudaSuccess) ? 0 : 1;
}

void launch_conv_tc(const TcMaps& maps, const TcConvArgs& a, cudaStream_t st) {
  ProfScope prof_scope__(PROF_CONV, a.alg_flops, st);
  dim3 grid((a.rows + 127) / 128, a.tasks);
  if (a.ncols == 64) conv_tc_kernel<64><<<grid, 192, tc_conv_smem_bytes(64), st>>>(maps, a);
  else if (a.ncols == 48) conv_tc_kernel<48><<<grid, 192, tc_conv_smem_bytes(48), st>>>(maps, a);
  else if (a.ncols == 32) conv_tc_kernel<32><<<grid, 192, tc_conv_smem_bytes(32), st>>>(maps, a);
  else conv_tc_kernel<16><<<grid, 192, tc_conv_smem_bytes(16), st>>>(maps, a);
  CUDA_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------
// weight packs: fast weights of blocks l >= 1 split into TF32 hi/lo, in both operand orientations
//   plane 0/1: W  [tap][c][f] hi/lo   (dgrad:  B[n = c][k = f])
//   plane 2/3: WT [tap][f][c] hi/lo   (conv:   B[n = f][k = c])
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void pack_weights_kernel(ParamLayout pl, const float* __restrict__ theta, long long theta_task_stride,
                                    float* __restrict__ pack, long long pack_task_stride, long long plane_stride) {
  const int task = blockIdx.y;
  const long long per_layer = 9LL * pl.F * pl.F;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_layer * (pl.L - 1)) return;
  const int l = 1 + (int)(i / per_layer);
  const long long rel = i - (long long)(l - 1) * per_layer;      // (tap, c, f)
  const int f = (int)(rel % pl.F);
  const int c = (int)((rel / pl.F) % pl.F);
  const int tap = (int)(rel / ((long long)pl.F * pl.F));
  const float x = theta[(long long)task * theta_task_stride + pl.w_off[l] + rel];
  const float hi = tf32_rna(x), lo = tf32_rna(x - hi);
  float* p = pack + (long long)task * pack_task_stride + (long long)(l - 1) * per_layer;
  p[rel] = hi;
  p[plane_stride + rel] = lo;
  const long long t = ((long long)tap * pl.F + f) * pl.F + c;
  p[2 * plane_stride + t] = hi;
  p[3 * plane_stride + t] = lo;
}

void launch_pack_weights(const ParamLayout& pl, const float* theta, long long theta_task_stride, float* pack,
                         long long pack_task_stride, long long plane_stride, int tasks, cudaStream_t st) {
  ProfScope prof_scope__(PROF_PARAM, 0.0, st);
  const long long n = 9LL * pl.F * pl.F * (pl.L - 1);
  if (n <= 0) return;
  dim3 grid((unsigned)((n + 255) / 256), tasks);
  pack_weights_kernel<<<grid, 256, 0, st>>>(pl, theta, theta_task_stride, pack, pack_task_stride, plane_stride);
  CUDA_CHECK_LAUNCH();
}
