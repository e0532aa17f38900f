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
// CTA = one 128-row M tile x all N (<= 64) columns.  Warp roles: warp 0 / warp 6 = TMA producers for B / A (one thread each),
// warp 1 = TMEM allocator + MMA issuer (one thread), warps 2..5 = epilogue (TMEM -> registers -> shared ->
// coalesced global store, + bias, + fp64 BatchNorm statistics).  4-stage smem ring, mbarrier full/empty,
// tcgen05.commit frees stages and publishes the accumulator.
#include <cuda.h>
#include <map>
#include <utility>
#include "common.cuh"
#include "tc_common.cuh"

namespace {


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
__device__ __forceinline__ void tc_mma_tf32_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
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

// thread-block-cluster helpers (split-K over the CTAs of one cluster, reduction through distributed shared memory)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsmem_st4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 dsmem_ld4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// Split-K reduction of one output tile across the S CTAs of a cluster.  CTA `zrank` finishes rows [zrank * 128/S, ...):
// every thread first issues ALL its distributed-shared-memory loads (S per float4 item, compile-time unrolled), then
// sums them in rank order (deterministic), adds the bias and writes global memory + a local copy for the statistics.
template <int NCOLS, int S>
__device__ __forceinline__ void splitk_reduce(uint32_t tile_local, int zrank, int et, const float* __restrict__ bias,
                                              float* __restrict__ tile2, float* __restrict__ out, int j0, int rows) {
  constexpr int P4 = NCOLS + 4;                         // partial-tile pitch (floats), keeps rows 16-byte aligned
  constexpr int ROWS = 128 / S;
  constexpr int TOTAL4 = ROWS * (NCOLS / 4);
  constexpr int ITEMS = (TOTAL4 + 127) / 128;
  float4 v[ITEMS][S];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int idx = et + it * 128;
    const int rr = idx / (NCOLS / 4), c4 = idx - rr * (NCOLS / 4);
    const uint32_t off = (uint32_t)(((zrank * ROWS + rr) * P4 + c4 * 4) * 4);
#pragma unroll
    for (int z = 0; z < S; ++z)
      v[it][z] = (idx < TOTAL4) ? dsmem_ld4(dsmem_addr(tile_local + off, (uint32_t)z)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int idx = et + it * 128;
    if (idx >= TOTAL4) continue;
    const int rr = idx / (NCOLS / 4), c4 = idx - rr * (NCOLS / 4);
    float4 acc = v[it][0];
#pragma unroll
    for (int z = 1; z < S; ++z) { acc.x += v[it][z].x; acc.y += v[it][z].y; acc.z += v[it][z].z; acc.w += v[it][z].w; }
    if (bias) { acc.x += bias[c4 * 4]; acc.y += bias[c4 * 4 + 1]; acc.z += bias[c4 * 4 + 2]; acc.w += bias[c4 * 4 + 3]; }
    *reinterpret_cast<float4*>(tile2 + rr * P4 + c4 * 4) = acc;
    const int g = j0 + zrank * ROWS + rr;
    if (g < rows) *reinterpret_cast<float4*>(out + (long long)g * NCOLS + c4 * 4) = acc;
  }
}


// Push-based variant (default): every CTA has already WRITTEN the rows it does not own into the owner's receive buffer
// recv[source rank][128 / S rows][NCOLS + 4] (st.shared::cluster, before the one cluster barrier), so the owner sums S
// LOCAL buffers -- no remote load latency, and nobody reads a peer's shared memory after the barrier, so the second
// cluster barrier ("do not exit while a peer still reads") is gone.  Same summation order as the pull variant.
template <int NCOLS, int S>
__device__ __forceinline__ void splitk_reduce_local(const float* __restrict__ recv, int zrank, int et, const float* __restrict__ bias,
                                                    float* __restrict__ tile2, float* __restrict__ out, int j0, int rows) {
  constexpr int P4 = NCOLS + 4;
  constexpr int ROWS = 128 / S;
  constexpr int TOTAL4 = ROWS * (NCOLS / 4);
  constexpr int ITEMS = (TOTAL4 + 127) / 128;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int idx = et + it * 128;
    if (idx >= TOTAL4) continue;
    const int rr = idx / (NCOLS / 4), c4 = idx - rr * (NCOLS / 4);
    float4 v[S];
#pragma unroll
    for (int z = 0; z < S; ++z) v[z] = *reinterpret_cast<const float4*>(recv + ((z * ROWS + rr) * P4 + c4 * 4));
    float4 acc = v[0];
#pragma unroll
    for (int z = 1; z < S; ++z) { acc.x += v[z].x; acc.y += v[z].y; acc.z += v[z].z; acc.w += v[z].w; }
    if (bias) { acc.x += bias[c4 * 4]; acc.y += bias[c4 * 4 + 1]; acc.z += bias[c4 * 4 + 2]; acc.w += bias[c4 * 4 + 3]; }
    *reinterpret_cast<float4*>(tile2 + rr * P4 + c4 * 4) = acc;
    const int g = j0 + zrank * ROWS + rr;
    if (g < rows) *reinterpret_cast<float4*>(out + (long long)g * NCOLS + c4 * 4) = acc;
  }
}

// K-major SWIZZLE_128B descriptor whose start is `row_off` rows into a 1024B-aligned tile.  Measured on B200: the
// 128B swizzle XOR is a function of the ABSOLUTE shared-memory address bits [7,10) (as for TMA writes), so a start
// address moved by row_off * 128 B addresses rows row_off .. row_off+127 of the tile correctly with the
// matrix-base-offset field left at 0; setting base_offset = row_off mod 8 (bo_mode = 1) gives wrong results.
// This is what lets ONE halo tile serve all nine filter taps.
// (the issuer below adds row_off * 128 B >> 4 to the descriptor's address field)

// debug timeline (clock64 at pipeline milestones of CTA (0,0)); written only when TcConvArgs::timeline != 0
__device__ long long g_tc_timeline[16];
#define TC_MARK(i) do { if (a.timeline && blockIdx.x == 0 && blockIdx.y == 0) g_tc_timeline[i] = clock64(); } while (0)

template <int NCOLS>
__global__ void __launch_bounds__(224, 1) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcConvArgs a) {
  pdl_trigger();
  trace_mark(22, a.tag);
  constexpr int B_BYTES = NCOLS * 128;
  constexpr int BSTAGE = 2 * B_BYTES;            // B_hi + B_lo of one (tap, k-chunk)
  // accumulators: plain mode 4 x hi*hi (round-robin over k-steps) + 1 x (lo*hi + hi*lo) = 5 * NCOLS columns;
  // stacked mode 4 x [hi*hi | hi*lo + lo*hi] = 8 * NCOLS columns (see the MMA issuer)
  constexpr int TMEM_COLS = NCOLS <= 16 ? 128 : (NCOLS <= 32 ? 256 : 512);   // power of two >= 8 * NCOLS
  constexpr int PITCH = NCOLS + 1;
  constexpr int MAXB = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_full[2], a_empty[2], b_full[MAXB], b_empty[MAXB], accum_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int row_ok[128];
  __shared__ double sred[2][NCOLS][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int task = blockIdx.y;
  const int j0 = blockIdx.x * 128;
  if (threadIdx.x == 0) TC_MARK(0);
  const int kchunks = (a.kc + 31) >> 5;          // a ragged last chunk (kc = 16 or 48) is zero-filled by TMA
  const int nph = a.nsrc * kchunks;              // A phases: (source, k-chunk); 9 taps each
  const int nb = a.nb;                           // B ring depth
  // split-K: the gridDim.z CTAs of a cluster share one output tile; CTA z accumulates stages [st_lo, st_hi) of the
  // nph * 9 (phase, tap) stages and the partial tiles are summed through distributed shared memory in the epilogue
  const int nsplit = (int)gridDim.z, zrank = (int)blockIdx.z;
  const int st_lo = zrank * (nph * 9) / nsplit, st_hi = (zrank + 1) * (nph * 9) / nsplit;
  const int ph_lo = st_lo / 9, ph_hi = (st_hi - 1) / 9;
  // push variant: a CTA writes into its peers' shared memory as soon as ITS accumulators are done, so every CTA of the
  // cluster must be known to have started by then: arrive here, wait right before the first remote store (free by then)
  if (nsplit > 1 && a.push) cluster_arrive_relaxed();
  const int abuf = a.rpad * 128;                 // bytes of one A halo buffer (hi or lo)
  uint8_t* bring = smem + 4 * (size_t)abuf;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
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
  pdl_wait();          // set-up above overlapped the previous kernel's tail; its results are needed from here on
  if (threadIdx.x == 0) TC_MARK(1);

  if (warp == 6) {
    if (lane == 0) {
      // ===== TMA producer A: per phase (source, k-chunk) ONE halo tile of A (hi, lo); all 9 taps read it at row
      // offsets.  Double-buffered, so phase ph+1 streams in while the MMAs of phase ph run.
      for (int ph = ph_lo; ph <= ph_hi; ++ph) {
        const int s = ph / kchunks;
        const int kc0 = (ph - s * kchunks) << 5;
        const int lp = ph - ph_lo;
        const int ab = lp & 1;
        mbar_wait(&a_empty[ab], (((uint32_t)lp >> 1) & 1u) ^ 1u);
        const int arow = a.a_row_base[s] + task * a.a_task_rows[s] + j0 - a.halo;
        const uint32_t ad = smem_u32(smem + (size_t)ab * 2 * abuf);
        mbar_arrive_expect_tx(&a_full[ab], 2u * (uint32_t)abuf);
        tma_load_2d(ad, &maps.m[s * 4 + 0], &a_full[ab], kc0, arow);
        tma_load_2d(ad + abuf, &maps.m[s * 4 + 1], &a_full[ab], kc0, arow);
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer B: one (B_hi, B_lo) stage per (phase, tap) through the ring
      int stage = 0; uint32_t bphase = 0;
      for (int st = st_lo; st < st_hi; ++st) {
        const int ph = st / 9, tap = st - ph * 9;
        const int s = ph / kchunks;
        const int kc0 = (ph - s * kchunks) << 5;
        mbar_wait(&b_empty[stage], bphase ^ 1u);
        const int brow = a.b_row_base[s] + task * a.b_task_rows[s] + tap * NCOLS;
        const uint32_t bd = smem_u32(bring + (size_t)stage * BSTAGE);
        mbar_arrive_expect_tx(&b_full[stage], BSTAGE);
        tma_load_2d(bd, &maps.m[s * 4 + 2], &b_full[stage], kc0, brow);
        tma_load_2d(bd + B_BYTES, &maps.m[s * 4 + 3], &b_full[stage], kc0, brow);
        if (++stage == nb) { stage = 0; bphase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major,
      // N>>3 at bit 17, M>>4 at bit 24
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // The issuing thread is the serial resource of this kernel: descriptors are built once per phase / stage and
      // advanced with 64-bit adds on the 16-byte-unit address field (K step: +32 B = +2; tap: row_off * 128 B).
      const uint64_t desc_base = make_desc_sw128(0);
      int stage = 0; uint32_t bphase = 0; int kstep = 0;
      uint64_t ahd = 0, ald = 0;
      int sgn = 1, ab = 0;
      for (int st = st_lo; st < st_hi; ++st) {
        const int ph = st / 9, tap = st - ph * 9;
        if (tap == 0 || st == st_lo) {
          const int lp = ph - ph_lo;
          ab = lp & 1;
          mbar_wait(&a_full[ab], ((uint32_t)lp >> 1) & 1u);
          if (st == st_lo) TC_MARK(2);
          const uint32_t a_hi = smem_u32(smem + (size_t)ab * 2 * abuf);
          ahd = desc_base + (uint64_t)(a_hi >> 4);
          ald = ahd + (uint64_t)(abuf >> 4);
          sgn = a.sign[ph / kchunks];
        }
        mbar_wait(&b_full[stage], bphase);
        if (st == st_lo) TC_MARK(3);
        if (st == st_lo + 9) TC_MARK(4);
        tc_fence_after();
        const int ty = tap / 3;
        const int row_off = a.halo + sgn * ((ty - 1) * a.gw + (tap - 3 * ty - 1));    // in [0, 2 * halo]
        const uint64_t ah0 = ahd + (uint64_t)(row_off * 8);
        const uint64_t al0 = ald + (uint64_t)(row_off * 8);
        const uint64_t bh0 = desc_base + (uint64_t)(smem_u32(bring + (size_t)stage * BSTAGE) >> 4);
        const uint64_t bl0 = bh0 + (uint64_t)(B_BYTES >> 4);
        if (a.stack) {
          // N-stacked 3xTF32: B_hi and B_lo of a stage are adjacent K-major tiles, so ONE descriptor with N = 2 * NCOLS
          // reads [B_hi; B_lo] and A_hi x [B_hi; B_lo]^T lands as [hi*hi | hi*lo] in 2 * NCOLS adjacent TMEM columns:
          // two instructions per k-step instead of three, and A_hi crosses the shared-memory port once instead of twice
          // (14 KB instead of 18 KB of operand fetch per k-step at N = 64 -- the measured floor of this kernel).
          // lo*hi is added to the small block.  Round-robin over 4 such column blocks keeps <= 18 sequential
          // accumulations per accumulator (the tensor core's fp32 accumulation truncates, see the header).
          constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * NCOLS) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
          const uint32_t first = (kstep == 0) ? 0u : 1u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma_tf32(tmem_base + (uint32_t)k * 2 * NCOLS, ah0 + 2 * k, bh0 + 2 * k, idesc2, first);
            tc_mma_tf32_acc(tmem_base + (uint32_t)k * 2 * NCOLS + NCOLS, al0 + 2 * k, bh0 + 2 * k, idesc);
          }
        } else if (kstep == 0) {
          // first four k-steps initialise the five accumulators
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma_tf32(tmem_base + 4 * NCOLS, al0 + 2 * k, bh0 + 2 * k, idesc, k > 0 ? 1u : 0u);
            tc_mma_tf32_acc(tmem_base + 4 * NCOLS, ah0 + 2 * k, bl0 + 2 * k, idesc);
            tc_mma_tf32(tmem_base + (uint32_t)k * NCOLS, ah0 + 2 * k, bh0 + 2 * k, idesc, 0u);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma_tf32_acc(tmem_base + 4 * NCOLS, al0 + 2 * k, bh0 + 2 * k, idesc);
            tc_mma_tf32_acc(tmem_base + 4 * NCOLS, ah0 + 2 * k, bl0 + 2 * k, idesc);
            tc_mma_tf32_acc(tmem_base + (uint32_t)k * NCOLS, ah0 + 2 * k, bh0 + 2 * k, idesc);
          }
        }
        kstep += 4;
        tc_commit(&b_empty[stage]);              // frees this B stage once the MMAs above have read it
        if (++stage == nb) { stage = 0; bphase ^= 1u; }
        if (tap == 8 || st == st_hi - 1) tc_commit(&a_empty[ab]);   // A halo buffer pair reusable after this phase's MMAs
      }
      TC_MARK(5);
      tc_commit(&accum_bar);                     // accumulators complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4.  Thread (q, lane) owns tile row r = 32 q + lane:
    // it sums the five accumulators, adds the bias and stores its 4*NCOLS contiguous bytes straight to global;
    // a shared-memory copy of the tile is kept only when BatchNorm statistics are wanted (column sums).
    const int et = threadIdx.x - 64;             // 0..127
    const int q = warp & 3;
    const int r = q * 32 + lane;
    __shared__ float s_bias[NCOLS];
    {
      const int row = j0 + et;
      int ok = 0;
      if (row < a.rows) {
        const int rr = row % a.G;
        const int yy = rr / a.gw, xx = rr - yy * a.gw;
        ok = (yy >= 1 && yy <= a.h && xx >= 1 && xx <= a.w) ? 1 : 0;
      }
      row_ok[et] = ok;
      if (et < NCOLS) s_bias[et] = a.bias ? a.bias[(long long)task * a.bias_stride + et] : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (a.zstage) {
      // tangent mode: the statistics below need the primal zh of the rows this CTA finishes.  Copy them into shared
      // memory NOW (coalesced float4 loads, a region of its own behind the ring / receive buffer) while the MMAs run --
      // the statistics loop then reads shared memory instead of one dependent global load per row (tangent-mode launches
      // ran 5.0 / 3.1 / 1.7 us longer than forward-mode launches of the same grid, ncu)
      constexpr int ZP = NCOLS + 4, Q = NCOLS / 4;
      float* zbuf = reinterpret_cast<float*>(bring + (size_t)nb * BSTAGE + ((nsplit > 1 && a.push) ? (size_t)128 * (NCOLS + 4) * 4 : 0));
      const float* zhg = a.zh + (long long)task * a.zh_stride;
      const int rows_own = 128 / nsplit, row0 = j0 + zrank * rows_own;
      for (int idx = et; idx < rows_own * Q; idx += 128) {
        const int rr = idx / Q, c4 = idx - rr * Q;
        const int gr = row0 + rr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < a.rows) v = *reinterpret_cast<const float4*>(zhg + (long long)gr * NCOLS + c4 * 4);
        *reinterpret_cast<float4*>(zbuf + rr * ZP + c4 * 4) = v;
      }
    }
    mbar_wait(&accum_bar, 0);
    if (et == 0) TC_MARK(6);
    tc_fence_after();
    float* tile = reinterpret_cast<float*>(smem);  // all MMAs have completed: operand buffers are free
    const bool want_stats = (a.mode != CONV_PLAIN);
    const int grow = j0 + r;
    float* orow = a.out + (long long)task * a.out_stride + (long long)grow * NCOLS;
    float* push_local = nullptr;
    uint32_t push_row = 0;                       // split-K, push variant: this thread's row inside the owner's receive buffer
    if (nsplit > 1 && a.push) {
      cluster_wait();                            // phase 1 (arrived at kernel start): all peers are running
      const int rows_per = 128 / nsplit, owner = r / rows_per, rloc = r - owner * rows_per;
      push_row = dsmem_addr(smem_u32(bring + (size_t)nb * BSTAGE) + (uint32_t)(((zrank * rows_per + rloc) * (NCOLS + 4)) * 4), (uint32_t)owner);
      if (owner == zrank) push_local = reinterpret_cast<float*>(bring + (size_t)nb * BSTAGE) + (zrank * rows_per + rloc) * (NCOLS + 4);   // own rows: plain st.shared
    }
#pragma unroll
    for (int c0 = 0; c0 < NCOLS; c0 += 16) {
      uint32_t v0[16], v1[16], v2[16], v3[16], v4[16];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      if (a.stack) {
        // blocks k = 0..3 at columns k * 2N: [hi*hi | small]; sum the four small blocks first (into v4), then load the big ones
        tc_ld16(ta + NCOLS, v0);
        tc_ld16(ta + 3 * NCOLS, v1);
        tc_ld16(ta + 5 * NCOLS, v2);
        tc_ld16(ta + 7 * NCOLS, v3);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          v4[i] = __float_as_uint((__uint_as_float(v0[i]) + __uint_as_float(v1[i])) + (__uint_as_float(v2[i]) + __uint_as_float(v3[i])));
        tc_ld16(ta, v0);
        tc_ld16(ta + 2 * NCOLS, v1);
        tc_ld16(ta + 4 * NCOLS, v2);
        tc_ld16(ta + 6 * NCOLS, v3);
      } else {
        tc_ld16(ta, v0);
        tc_ld16(ta + NCOLS, v1);
        tc_ld16(ta + 2 * NCOLS, v2);
        tc_ld16(ta + 3 * NCOLS, v3);
        tc_ld16(ta + 4 * NCOLS, v4);
      }
      tc_wait_ld();
      float o[16];
      if (nsplit == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float big = (__uint_as_float(v0[i]) + __uint_as_float(v1[i])) + (__uint_as_float(v2[i]) + __uint_as_float(v3[i]));
          o[i] = (big + __uint_as_float(v4[i])) + s_bias[c0 + i];
        }
        if (grow < a.rows) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(orow + c0 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        }
        if (want_stats) {
#pragma unroll
          for (int i = 0; i < 16; ++i) tile[r * PITCH + c0 + i] = o[i];
        }
      } else {
        // split-K partial (no bias): parked in shared memory (16-byte aligned rows) for the cluster reduction below
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float big = (__uint_as_float(v0[i]) + __uint_as_float(v1[i])) + (__uint_as_float(v2[i]) + __uint_as_float(v3[i]));
          o[i] = big + __uint_as_float(v4[i]);
        }
        if (a.push) {
          // row r of the tile belongs to cluster rank r / (128 / nsplit): straight into that CTA's receive buffer
          if (push_local != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(push_local + c0 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; i += 4) dsmem_st4(push_row + (uint32_t)((c0 + i) * 4), make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(tile + r * (NCOLS + 4) + c0 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        }
      }
    }
    if (et == 0) TC_MARK(7);
  }

  // rows [row_lo, row_lo + row_n) of the tile are finished by this CTA (all 128 without split-K)
  int row_lo = 0, row_n = 128, spitch = PITCH;
  const float* sbuf = reinterpret_cast<const float*>(smem);   // where those rows live (row index relative to row_lo)
  if (nsplit > 1) {
    __syncwarp();
    if (a.push && !(warp >= 2 && warp < 6)) cluster_wait();     // phase 1 for the non-epilogue warps
    cluster_sync_all();                            // pull: every CTA's partial tile is in its shared memory; push: every receive buffer is complete
    if (threadIdx.x == 64) TC_MARK(10);
    row_n = 128 / nsplit; row_lo = zrank * row_n; spitch = NCOLS + 4;
    float* tile2 = reinterpret_cast<float*>(smem + 40 * 1024);
    sbuf = tile2;
    if (warp >= 2 && warp < 6) {
      const int et = threadIdx.x - 64;
      const uint32_t tile_local = smem_u32(smem);
      const float* bias = a.bias ? a.bias + (long long)task * a.bias_stride : nullptr;
      float* outp = a.out + (long long)task * a.out_stride;
      if (a.push) {
        const float* recv = reinterpret_cast<const float*>(bring + (size_t)nb * BSTAGE);
        if (nsplit == 2) splitk_reduce_local<NCOLS, 2>(recv, zrank, et, bias, tile2, outp, j0, a.rows);
        else if (nsplit == 4) splitk_reduce_local<NCOLS, 4>(recv, zrank, et, bias, tile2, outp, j0, a.rows);
        else splitk_reduce_local<NCOLS, 8>(recv, zrank, et, bias, tile2, outp, j0, a.rows);
      } else if (nsplit == 2) splitk_reduce<NCOLS, 2>(tile_local, zrank, et, bias, tile2, outp, j0, a.rows);
      else if (nsplit == 4) splitk_reduce<NCOLS, 4>(tile_local, zrank, et, bias, tile2, outp, j0, a.rows);
      else splitk_reduce<NCOLS, 8>(tile_local, zrank, et, bias, tile2, outp, j0, a.rows);
    }
  }

  if (threadIdx.x == 64) TC_MARK(11);
  if (warp >= 2 && warp < 6) {
    const int et = threadIdx.x - 64;
    const bool want_stats = (a.mode != CONV_PLAIN);
    if (want_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float* zh = a.zh ? a.zh + (long long)task * a.zh_stride : nullptr;
      const float* zbuf_s = reinterpret_cast<const float*>(bring + (size_t)nb * BSTAGE + ((nsplit > 1 && a.push) ? (size_t)128 * (NCOLS + 4) * 4 : 0));
      constexpr int PARTS = 128 / NCOLS;           // 2 for 64 and 48, 4 for 32, 8 for 16
      const int col = et % NCOLS, part = et / NCOLS;
      double s1 = 0.0, s2 = 0.0;
      if (part < PARTS) {
        // (a 4-row-batched variant of this loop -- all loads of a batch issued before the first add, four independent fp64
        // chains -- and an integer re-bias instead of F2F.F64.F32 both measured SLOWER: 11.6 -> 12.0 ms on Mini-ImageNet,
        // 19.1 -> 19.7 / 20.0 ms on 20-way; see DESIGN.md, negative results)
        for (int rr = part; rr < row_n; rr += PARTS) {
          if (row_ok[row_lo + rr]) {
            const float v = sbuf[rr * spitch + col];
            if (a.mode == CONV_FWD_STATS) { s1 += (double)v; s2 += (double)v * (double)v; }
            else {
              const float zv = a.zstage ? zbuf_s[rr * (NCOLS + 4) + col] : zh[(long long)(j0 + row_lo + rr) * NCOLS + col];
              s1 += (double)v; s2 += (double)zv * (double)v;
            }
          }
        }
        if (part < 2) { sred[part][col][0] = s1; sred[part][col][1] = s2; }
      }
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
  if (threadIdx.x == 64) TC_MARK(12);
  if (nsplit > 1 && !a.push) {
    __syncwarp();
    cluster_sync_all();                            // pull variant: nobody leaves while a peer still reads its partial tile
  }

  if (threadIdx.x == 64) TC_MARK(8);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_MARK(9);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
  trace_mark(22 | 0x80, a.tag);     // end of CTA (0,0,0)
}

}  // namespace

int tc_read_timeline(long long* out16) { return cudaMemcpyFromSymbol(out16, g_tc_timeline, 16 * sizeof(long long)) == cudaSuccess ? 0 : 1; }

// halo tile rows (multiple of 8) for a grid of pitch gw, and the deepest B ring that fits next to 4 halo buffers
int tc_conv_rpad(int gw) { return ((128 + 2 * (gw + 1)) + 7) / 8 * 8; }
static int g_tc_ring_cap = 8;         // env MAML_B200_TC_NB: B ring depth cap (a shallower ring leaves shared memory to co-resident kernels)
void tc_conv_set_ring_cap(int nb) { g_tc_ring_cap = nb < 2 ? 2 : (nb > 8 ? 8 : nb); }
int tc_conv_ring(int ncols, int gw) {
  const long long avail = 227LL * 1024 - 4096 /* static smem */ - 1024 /* alignment */ - 4LL * tc_conv_rpad(gw) * 128;
  long long nb = avail / (2LL * ncols * 128);
  if (nb > g_tc_ring_cap) nb = g_tc_ring_cap;
  return (int)nb;
}
size_t tc_conv_smem_bytes(int ncols, int gw) {
  return (size_t)4 * tc_conv_rpad(gw) * 128 + (size_t)tc_conv_ring(ncols, gw) * 2 * ncols * 128 + 1024;
}
static size_t tc_conv_smem_for(int ncols, int gw, int nb) { return (size_t)4 * tc_conv_rpad(gw) * 128 + (size_t)nb * 2 * ncols * 128 + 1024; }

int tc_conv_prepare() {
  const int maxs = 227 * 1024 - 4096;    // static shared memory (barriers, row flags, fp64 partials) takes the rest
  cudaError_t e1 = cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs);
  cudaError_t e2 = cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs);
  cudaError_t e3 = cudaFuncSetAttribute(conv_tc_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs);
  cudaError_t e4 = cudaFuncSetAttribute(conv_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs);
  return (e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess && e4 == cudaSuccess) ? 0 : 1;
}

// Split-K factor: small layers have fewer tiles than SMs (Omniglot block 3: 8 tiles, block 2: 32), and one tile's
// serial pipeline (~18 stages x ~1000 cycles) is then the whole kernel.  Spreading the (phase, tap) stages of a tile
// over a cluster of S CTAs shortens that to 18 / S stages + one distributed-shared-memory reduction.  S is the largest
// of {8, 4, 2} whose clusters are all co-resident (asked from the occupancy calculator once per shape).
static int g_tc_zstage = 1;          // env MAML_B200_TC_ZSTAGE=0: tangent-mode statistics read the primal zh from global memory row by row
void tc_conv_set_zstage(int on) { g_tc_zstage = on; }
static int g_tc_push = 1;            // env MAML_B200_TC_PUSH=0: pull-based split-K reduction (two cluster barriers)
void tc_conv_set_push(int on) { g_tc_push = on; }
static int g_tc_ring_fit = 1;        // env MAML_B200_TC_NB_FIT=0: keep the full ring for short pipelines
void tc_conv_set_ring_fit(int on) { g_tc_ring_fit = on; }
static int g_tc_split_max = 8;       // env MAML_B200_TC_SPLIT (1 disables split-K)
template <int NCOLS>
static int max_clusters(size_t smem, int S) {
  static std::map<std::pair<size_t, int>, int> cache;
  auto it = cache.find({smem, S});
  if (it != cache.end()) return it->second;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1, 1, S); cfg.blockDim = dim3(224); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = S;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<NCOLS>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[{smem, S}] = n;
  return n;
}

template <int NCOLS>
static void launch_conv_tc_n(const TcMaps& maps, const TcConvArgs& a_in, size_t smem, cudaStream_t st) {
  TcConvArgs a = a_in;
  const int tiles = ((a.rows + 127) / 128) * (a.plan_tasks > a.tasks ? a.plan_tasks : a.tasks);
  const int stages = a.nsrc * ((a.kc + 31) / 32) * 9;
  int S = 1;
  // push-based split-K epilogue: a receive buffer [128 rows][NCOLS + 4] behind the B ring (peers write it while this CTA's
  // MMAs may still read the operand buffers, so it cannot alias them); the ring gives up the stages that no longer fit
  const size_t recv_bytes = (size_t)128 * (NCOLS + 4) * 4;
  int nb_push = a.nb;
  {
    const long long avail = 227LL * 1024 - 4096 - 1024 - 4LL * tc_conv_rpad(a.gw) * 128 - (long long)recv_bytes;
    const long long fit = avail / (2LL * NCOLS * 128);
    if (fit < nb_push) nb_push = (int)fit;
  }
  const bool push = g_tc_push && nb_push >= 2;
  if (push) smem = tc_conv_smem_for(NCOLS, a.gw, nb_push) + recv_bytes;
  int smax = g_tc_split_max;
  if (a.split_cap > 0 && a.split_cap < smax) { smax = 1; while (smax * 2 <= a.split_cap) smax *= 2; }
  for (int cand = smax; cand >= 2; cand >>= 1) {
    if (cand > 8 || stages < 2 * cand) continue;
    if (tiles <= max_clusters<NCOLS>(smem, cand)) { S = cand; break; }
  }
  dim3 grid((a.rows + 127) / 128, a.tasks, S);
  // a CTA never has more than ceil(stages / S) B stages in flight: a ring deeper than that only takes shared memory
  // away from the kernels of the other streams that could share the SM (block-0 / BatchNorm kernels need 10-27 KB)
  if (g_tc_ring_fit) {
    const int per_cta = (stages + S - 1) / S;
    if (a.nb > per_cta) a.nb = per_cta < 2 ? 2 : per_cta;
  }
  a.push = (S > 1 && push) ? 1 : 0;
  if (a.push) a.nb = nb_push;
  // tangent mode: room for the primal zh rows this CTA finishes (see the kernel); the ring gives up stages if it must
  a.zstage = 0;
  size_t zbytes = 0;
  if (g_tc_zstage && a.mode == CONV_TAN_STATS && a.zh != nullptr) {
    zbytes = (size_t)(128 / S) * (NCOLS + 4) * 4;
    const long long limit = 227LL * 1024 - 4096;
    int nb2 = a.nb;
    auto total = [&](int nbx) { return (long long)tc_conv_smem_for(NCOLS, a.gw, nbx) + (long long)(a.push ? recv_bytes : 0) + (long long)zbytes; };
    while (nb2 > 2 && total(nb2) > limit) --nb2;
    if (total(nb2) <= limit) { a.nb = nb2; a.zstage = 1; } else zbytes = 0;
  }
  smem = tc_conv_smem_for(NCOLS, a.gw, a.nb) + (a.push ? recv_bytes : 0) + zbytes;
  if (S == 1) {
    launch_pdl(conv_tc_kernel<NCOLS>, grid, dim3(224), smem, st, maps, tagged(a));
    return;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(224); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = S;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = ((g_pdl_cluster & 2) && pdl_allowed(st)) ? 2 : 1;
  cudaLaunchKernelEx(&cfg, conv_tc_kernel<NCOLS>, maps, tagged(a));
}

void tc_conv_set_split(int max_split) { g_tc_split_max = max_split < 1 ? 1 : (max_split > 8 ? 8 : max_split); }

void launch_conv_tc(const TcMaps& maps, const TcConvArgs& a, cudaStream_t st) {
  ProfScope prof_scope__(PROF_CONV, a.alg_flops, st);
  const size_t smem = tc_conv_smem_bytes(a.ncols, a.gw);
  if (a.ncols == 64) launch_conv_tc_n<64>(maps, a, smem, st);
  else if (a.ncols == 48) launch_conv_tc_n<48>(maps, a, smem, st);
  else if (a.ncols == 32) launch_conv_tc_n<32>(maps, a, smem, st);
  else launch_conv_tc_n<16>(maps, a, smem, st);
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
                                    float* __restrict__ pack, long long pack_task_stride, long long plane_stride, int tag) {
  pdl_prologue(21, tag);
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
  launch_pdl(pack_weights_kernel, dim3(grid), dim3(256), (size_t)(0), st, pl, theta, theta_task_stride, pack, pack_task_stride, plane_stride, launch_tag());
  CUDA_CHECK_LAUNCH();
}

MAML_TRACE_SETTER(trace_set_tc)
