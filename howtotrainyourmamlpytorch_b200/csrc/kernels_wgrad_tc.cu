// tcgen05 / TMA weight gradient of the 3x3 convolutions of blocks l >= 1, fp32-faithful via the 3xTF32 operand split.
//
//   dW[tap][c][f] = sum_src sum_j  A_src[j + s_tap, c] * D_src[j, f]            (+ db[f] = sum_j D_0[j, f])
//
// (autograd of reference meta_neural_network_architectures.py:89-97 w.r.t. the weight; two operand pairs in the
// Hessian-vector pass).  The reduction runs over PIXELS, and both operands are stored pixel-major ([grid row][channel],
// common.cuh) -- i.e. MN-major for this GEMM (M = c, N = f, K = j).  tcgen05.mma.kind::tf32 takes MN-major operands from
// shared memory in exactly one layout (measured with scripts/umma_mn32b_probe.cu, and stated by CUTLASS' sm100 builder):
// TMA swizzle 128B_ATOM_32B + descriptor layout type 1 (SWIZZLE_128B_BASE32B): a tile is [K rows][32 floats] with 128 B
// rows, atoms of 32 channels follow each other at the leading byte offset, groups of 4 K-rows at the stride byte offset
// (512 B).  So the hi / lo planes the BatchNorm kernels already write for the conv kernel feed this kernel unchanged --
// no transposed copies (round 1 concluded the opposite from a probe that only tried layout type 2).
//
// A filter tap is a ROW shift of A (s_tap = (ky-1) gw + (kx-1)), i.e. a shift along K: one [34 x 32] tile per stage
// serves the three taps of a filter row through descriptors whose start address is moved by kx * 128 B (the swizzle is
// a function of the absolute shared-memory address, measured), the ky shift is in the TMA coordinate.
// CTA = (row chunk, filter row ky, task).  Per tap and K step TWO M = 128, N = 64 instructions: the A descriptor spans
// the four atoms [A_hi(c 0..31), A_hi(c 32..63), A_lo(..), A_lo(..)], so rows 0..63 of the accumulator get hi*hi (then
// hi*lo from the second instruction, B = D_lo) and rows 64..127 get lo*hi (+ lo*lo, 2^-22, harmless); dW = rows c +
// rows c + 64.  The dead conv-bias gradient (column sums of D) rides along as an all-ones A tile in the ky = 1 CTAs.
// Precision: the tensor core's fp32 accumulation truncates, so no accumulator may take a long chain of accumulations
// (one accumulator per 512 rows measured 4x the FFMA path's error on Mini-ImageNet).  Every 8 K-steps (64 rows, 16
// accumulations) the issuer switches to the other of two TMEM accumulator sets and the epilogue warps drain the finished
// one into fp32 REGISTER accumulators with IEEE adds while the tensor core fills the other (a first version that kept
// them in shared memory ran slower than the FFMA kernel: its read-modify-writes competed with the MMA operand fetch for
// the shared-memory port) -- chunk length is then a pure scheduling choice.  Chunks are summed in fp32 by param_reduce
// in fixed order (deterministic).
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int WG_STAGE_ROWS = 32;                 // K rows per stage (4 MMA K-steps of 8)
constexpr int WG_A_ROWS = 40;                     // A tile rows: 32 + 2 (three kx taps) padded so that tiles stay 1 KB aligned
constexpr int WG_A_TILE = WG_A_ROWS * 128;        // bytes of one 32-channel atom of A
constexpr int WG_B_TILE = WG_STAGE_ROWS * 128;
constexpr int WG_STAGE_BYTES = 4 * WG_A_TILE + 4 * WG_B_TILE;     // A_hi0 A_hi1 A_lo0 A_lo1 | D_hi0 D_hi1 D_lo0 D_lo1
constexpr int WG_NSTAGE = 4;                      // deepest stage ring (WgTcArgs::nstage in [2, 4] is used at run time)
constexpr int WG_ONES_BYTES = 4096;               // 4 atoms x 8 rows x 128 B of 1.0f
constexpr int WG_SEG_KSTEPS = 8;                  // K-steps (of 8 rows) per accumulator segment
constexpr int WG_ACC_PITCH = 65;                  // final exchange buffer [tap][f][c] (reuses the stage ring), pitch 65: conflict-free
constexpr int WG_XCHG_BYTES = 3 * 64 * WG_ACC_PITCH * 4;
static_assert(WG_XCHG_BYTES <= 2 * WG_STAGE_BYTES, "exchange buffer must fit in the stage ring");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
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
// MN-major, SWIZZLE_128B_BASE32B shared-memory matrix descriptor: start >> 4 | LBO >> 4 at bit 16 (stride between the
// 32-float atoms along M / N) | SBO >> 4 at bit 32 (stride between groups of 4 K-rows = 512 B) | version 1 | layout 1
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// grid (nchunks, 3, tasks); 192 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// LITE (chunks of at most 48 K-steps over all sources -- the small Omniglot layers): no draining; the K-steps alternate
// between the two accumulator sets (<= 48 accumulations each) and the epilogue adds them once at the end.  It needs ~90
// registers per thread instead of 254, which matters because these CTAs share SMs with the latency-critical main chain.
template <bool LITE>
__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ WgTcArgs a) {
  pdl_trigger();
  trace_mark(27, a.tag);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[WG_NSTAGE], empty[WG_NSTAGE], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, ky = blockIdx.y, task = blockIdx.z;
  const int r_begin = chunk * a.rows_per_chunk;
  const int r_end = min(a.rows, r_begin + a.rows_per_chunk);
  const int nst = (r_end - r_begin + WG_STAGE_ROWS - 1) / WG_STAGE_ROWS;      // stages per source (>= 1: chunks are never empty)
  const int total = nst * a.nsrc;
  const int ks_src = (r_end - r_begin + 7) >> 3;                              // K-steps per source
  const int nseg = (ks_src * a.nsrc + WG_SEG_KSTEPS - 1) / WG_SEG_KSTEPS;    // accumulator segments
  const int nstg = a.nstage;                        // stage ring depth (2..WG_NSTAGE)
  uint8_t* ones = smem + (size_t)nstg * WG_STAGE_BYTES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstg; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    // all-ones A tile (4 atoms x 8 K-rows; constant data is invariant under the swizzle)
    float4* o4 = reinterpret_cast<float4*>(ones);
    for (int i = threadIdx.x - 64; i < WG_ONES_BYTES / 16; i += 128) o4[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the UMMA (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: per stage 4 A boxes ([40 rows][32 ch]: hi / lo x channel atoms 0 / 1) + 4 D boxes ([32][32])
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < total; ++it) {
        const int s = it / nst, i = it - s * nst;
        const int j = r_begin + i * WG_STAGE_ROWS;
        mbar_wait(&empty[stage], phase ^ 1u);
        const uint32_t base = smem_u32(smem + (size_t)stage * WG_STAGE_BYTES);
        const int arow = a.a_row_base[s] + task * a.a_task_rows[s] + j + (ky - 1) * a.gw - 1;
        const int brow = a.b_row_base[s] + task * a.b_task_rows[s] + j;
        mbar_arrive_expect_tx(&full[stage], WG_STAGE_BYTES);
        tma_load_2d(base + 0 * WG_A_TILE, &maps.m[s * 4 + 0], &full[stage], 0, arow);
        tma_load_2d(base + 1 * WG_A_TILE, &maps.m[s * 4 + 0], &full[stage], 32, arow);
        tma_load_2d(base + 2 * WG_A_TILE, &maps.m[s * 4 + 1], &full[stage], 0, arow);
        tma_load_2d(base + 3 * WG_A_TILE, &maps.m[s * 4 + 1], &full[stage], 32, arow);
        const uint32_t bb = base + 4 * WG_A_TILE;
        tma_load_2d(bb + 0 * WG_B_TILE, &maps.m[s * 4 + 2], &full[stage], 0, brow);
        tma_load_2d(bb + 1 * WG_B_TILE, &maps.m[s * 4 + 2], &full[stage], 32, brow);
        tma_load_2d(bb + 2 * WG_B_TILE, &maps.m[s * 4 + 3], &full[stage], 0, brow);
        tma_load_2d(bb + 3 * WG_B_TILE, &maps.m[s * 4 + 3], &full[stage], 32, brow);
        if (++stage == nstg) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer.  Instruction descriptor: D = F32, A = B = TF32, both MN-major (bits 15, 16), M = 128, N = 64.
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                                 ((uint32_t)(128 >> 4) << 24);
      const uint64_t ones_desc = make_desc_mn(smem_u32(ones), 1024u);
      int stage = 0; uint32_t phase = 0;
      int seg = 0, kcount = 0;
      uint32_t bias_started = 0;
      for (int it = 0; it < total; ++it) {
        const int s = it / nst, i = it - s * nst;
        const int j = r_begin + i * WG_STAGE_ROWS;
        const int ksteps = min(4, (r_end - j + 7) >> 3);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (size_t)stage * WG_STAGE_BYTES);
        const uint64_t ad = make_desc_mn(base, WG_A_TILE);                                    // atoms hi0 hi1 lo0 lo1: M = 128
        const uint64_t bh = make_desc_mn(base + 4 * WG_A_TILE, WG_B_TILE);                    // D_hi atoms 0, 1: N = 64
        const uint64_t bl = make_desc_mn(base + 4 * WG_A_TILE + 2 * WG_B_TILE, WG_B_TILE);    // D_lo
        for (int t = 0; t < ksteps; ++t) {
          int p; uint32_t acc;
          if constexpr (LITE) {
            p = kcount & 1; acc = kcount >= 2 ? 1u : 0u;      // K-steps alternate between the two sets; the first two initialise them
          } else {
            p = seg & 1;
            if (kcount == 0) {                     // first K-step of a segment: the epilogue must have drained this set
              mbar_wait(&acc_empty[p], (((uint32_t)seg >> 1) & 1u) ^ 1u);
              tc_fence_after();
            }
            acc = kcount > 0 ? 1u : 0u;
          }
          const uint64_t bo = (uint64_t)(t * 64);                                             // 8 rows x 128 B in 16-byte units
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint64_t ao = (uint64_t)((kx + 8 * t) * 8);                                 // (kx + 8 t) rows x 128 B
            const uint32_t d = tmem_base + (uint32_t)(p * 192 + kx * 64);
            tc_mma_tf32(d, ad + ao, bh + bo, idesc, acc);          // rows 0..63: hi*hi      rows 64..127: lo*hi
            tc_mma_tf32(d, ad + ao, bl + bo, idesc, 1u);           //             + hi*lo                  + lo*lo
          }
          if (ky == 1 && s == 0) {                 // db[f] = sum_j D_0[j, f] (every row of the tile); never drained: dead parameter
            tc_mma_tf32(tmem_base + 384u, ones_desc, bh + bo, idesc, bias_started);
            tc_mma_tf32(tmem_base + 384u, ones_desc, bl + bo, idesc, 1u);
            bias_started = 1u;
          }
          ++kcount;
          if constexpr (!LITE) {
            if (kcount == WG_SEG_KSTEPS) { tc_commit(&acc_full[p]); ++seg; kcount = 0; }
          }
        }
        tc_commit(&empty[stage]);
        if (++stage == nstg) { stage = 0; phase ^= 1u; }
      }
      if constexpr (LITE) tc_commit(&acc_full[0]);
      else if (kcount > 0) tc_commit(&acc_full[seg & 1]);
    }
  } else {
    // ===== epilogue warps: drain every finished accumulator segment into fp32 register accumulators (IEEE adds).
    // TMEM lane m (quadrant q = warp % 4, m = 32 q + lane) holds row m of the M = 128 tile: rows 0..63 = (hi*hi + hi*lo)
    // of channel c = m, rows 64..127 = (lo*hi + lo*lo) of channel c = m - 64.
    const int q = warp & 3;
    const int m = q * 32 + lane;
    float* xch = reinterpret_cast<float*>(smem);
    const int c = m & 63;
    float* P = a.partial + (long long)task * a.partial_task_stride + (long long)chunk * a.chunk_stride;
    const int CF = a.kc * a.ncols;
    if constexpr (LITE) {
      const bool two = (ks_src * a.nsrc) >= 2;                 // the second set exists only if a second K-step ran
      mbar_wait(&acc_full[0], 0);
      tc_fence_after();
      // pass 1: rows 64..127 (lo*hi + lo*lo) -> shared memory; pass 2: rows 0..63 add them and write dW
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const bool mine = (pass == 0) ? (m >= 64) : (m < 64);
#pragma unroll 1
        for (int kx = 0; kx < 3; ++kx) {
          float* prow = P + (long long)(ky * 3 + kx) * CF + (long long)c * a.ncols;
          for (int c0 = 0; c0 < a.ncols; c0 += 16) {
            uint32_t v0[16], v1[16];
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kx * 64 + c0);
            tc_ld16(ta, v0);
            if (two) tc_ld16(ta + 192u, v1);
            tc_wait_ld();
            if (!mine) continue;
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(v0[i]) + (two ? __uint_as_float(v1[i]) : 0.f);
            if (pass == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) xch[(kx * 64 + c0 + i) * WG_ACC_PITCH + c] = o[i];
            } else if (c < a.kc) {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(prow + c0 + i) =
                    make_float4(o[i] + xch[(kx * 64 + c0 + i) * WG_ACC_PITCH + c], o[i + 1] + xch[(kx * 64 + c0 + i + 1) * WG_ACC_PITCH + c],
                                o[i + 2] + xch[(kx * 64 + c0 + i + 2) * WG_ACC_PITCH + c], o[i + 3] + xch[(kx * 64 + c0 + i + 3) * WG_ACC_PITCH + c]);
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    } else {
      float acc[3][64];
  #pragma unroll
      for (int kx = 0; kx < 3; ++kx)
  #pragma unroll
        for (int i = 0; i < 64; ++i) acc[kx][i] = 0.f;
      for (int g = 0; g < nseg; ++g) {
        const int p = g & 1;
        mbar_wait(&acc_full[p], ((uint32_t)g >> 1) & 1u);
        tc_fence_after();
  #pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
  #pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t v[16];
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * 192 + kx * 64 + c0), v);
            tc_wait_ld();
  #pragma unroll
            for (int i = 0; i < 16; ++i) acc[kx][c0 + i] += __uint_as_float(v[i]);
          }
        }
        // every read of this set has completed (wait::ld): hand it back to the issuer
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[p])) : "memory");
      }
      // dW[c] = rows c + rows c + 64: the upper half goes through shared memory (the stage ring is free: every MMA has
      // completed), the lower half adds it and writes its row of each tap (F contiguous floats) to the partial buffer
      if (m >= 64) {
  #pragma unroll
        for (int kx = 0; kx < 3; ++kx)
  #pragma unroll
          for (int i = 0; i < 64; ++i) xch[(kx * 64 + i) * WG_ACC_PITCH + c] = acc[kx][i];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (m < 64 && c < a.kc) {
  #pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float* prow = P + (long long)(ky * 3 + kx) * CF + (long long)c * a.ncols;
  #pragma unroll
          for (int i = 0; i < 64; i += 4) {
            if (i < a.ncols) {
              float4 o;
              o.x = acc[kx][i] + xch[(kx * 64 + i) * WG_ACC_PITCH + c];
              o.y = acc[kx][i + 1] + xch[(kx * 64 + i + 1) * WG_ACC_PITCH + c];
              o.z = acc[kx][i + 2] + xch[(kx * 64 + i + 2) * WG_ACC_PITCH + c];
              o.w = acc[kx][i + 3] + xch[(kx * 64 + i + 3) * WG_ACC_PITCH + c];
              *reinterpret_cast<float4*>(prow + i) = o;
            }
          }
        }
      }
    }
    if (ky == 1 && q == 0) {
      float* pb = P + 9LL * CF;
      for (int c0 = 0; c0 < a.ncols; c0 += 16) {
        uint32_t v0[16];
        tc_ld16(tmem_base + 384u + (uint32_t)c0, v0);
        tc_wait_ld();
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pb[c0 + i] = __uint_as_float(v0[i]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  trace_mark(27 | 0x80, a.tag);
}

}  // namespace

static int g_wg_nstage = WG_NSTAGE;     // env MAML_B200_WG_NSTAGE: a shallower ring leaves shared memory to co-resident kernels
void wgrad_tc_set_stages(int n) { g_wg_nstage = n < 2 ? 2 : (n > WG_NSTAGE ? WG_NSTAGE : n); }
size_t wgrad_tc_smem_bytes() { return (size_t)WG_NSTAGE * WG_STAGE_BYTES + WG_ONES_BYTES + 1024; }
static size_t wgrad_tc_smem_for(int nstage) { return (size_t)nstage * WG_STAGE_BYTES + WG_ONES_BYTES + 1024; }

int wgrad_tc_prepare() {
  return (cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_tc_smem_bytes()) == cudaSuccess &&
          cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_tc_smem_bytes()) == cudaSuccess) ? 0 : 1;
}

void launch_wgrad_tc(const TcMaps& maps, const WgTcArgs& a, cudaStream_t st) {
  ProfScope prof_scope__(PROF_WGRAD, a.alg_flops, st);
  dim3 grid(a.nchunks, 3, a.tasks);
  const int ksteps = a.nsrc * ((a.rows_per_chunk + 7) / 8);
  WgTcArgs b = tagged(a);
  if (b.nstage < 2 || b.nstage > WG_NSTAGE) b.nstage = g_wg_nstage;
  if (ksteps <= 48 && !a.force_flush) launch_pdl(wgrad_tc_kernel<true>, grid, dim3(192), wgrad_tc_smem_for(b.nstage), st, maps, b);
  else launch_pdl(wgrad_tc_kernel<false>, grid, dim3(192), wgrad_tc_smem_for(b.nstage), st, maps, b);
  CUDA_CHECK_LAUNCH();
}

MAML_TRACE_SETTER(trace_set_wgtc)
