// Classifier-head device code shared by kernels_head.cu (stand-alone kernel) and kernels_bn.cu (fused last-block kernels).
#pragma once
#include "common.cuh"

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Body of the head for (row group `group`, task `task`): everything except the weight gradient is row-local, so each
// CTA owns `rows_per_cta` rows of the batch and writes its own chunk of (gW, gb); the parameter-space kernel sums the
// chunks in order.  `smh`: 5 * rows_per_cta * N floats of shared memory; 256 threads.  Also called by the fused
// last-block kernels (kernels_bn.cu).
__device__ __forceinline__ void head_body(const HeadArgs& a, int task, int group, float* smh, float* s_rowloss, float* s_rowcorrect) {
  const int n = a.n, N = a.N, D = a.D;
  const int row0 = group * a.rows_per_cta;
  const int nl = min(a.rows_per_cta, n - row0);          // local rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.rows_per_cta;
  float* logits = smh;
  float* prob = smh + R * N;
  float* dl = smh + 2 * R * N;
  float* ldot = smh + 3 * R * N;
  float* dldot = smh + 4 * R * N;
  const bool tan = (a.mode == HEAD_TANGENT);

  const float* f = a.f + (long long)task * a.f_stride + (long long)row0 * D;
  const float* W = a.Wfc + (long long)task * a.theta_stride;
  const float* b = a.bfc + (long long)task * a.theta_stride;
  const float* fd = tan ? a.fdot + (long long)task * a.fdot_stride + (long long)row0 * D : nullptr;
  const float* uW = tan ? a.uW + (long long)task * a.u_stride : nullptr;
  const float* ub = tan ? a.ub + (long long)task * a.u_stride : nullptr;
  const long long* y = a.y + (long long)task * a.y_stride + row0;

  for (int o = warp; o < nl * N; o += 8) {
    const int i = o / N, k = o - i * N;
    float s = 0.f, sd = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float fv = f[(long long)i * D + d], wv = W[(long long)k * D + d];
      s = fmaf(fv, wv, s);
      if (tan) sd += fd[(long long)i * D + d] * wv + fv * uW[(long long)k * D + d];
    }
    s = warp_sum(s);
    if (tan) sd = warp_sum(sd);
    if (lane == 0) {
      logits[o] = s + b[k];
      if (tan) ldot[o] = sd + ub[k];
    }
  }
  __syncthreads();

  const float wscale = a.scale;
  const float inv_n = 1.f / (float)n;
  for (int i = tid; i < nl; i += 256) {
    float mx = logits[i * N];
    int am = 0;
    for (int k = 1; k < N; ++k) {
      const float v = logits[i * N + k];
      if (v > mx) { mx = v; am = k; }
    }
    float se = 0.f;
    for (int k = 0; k < N; ++k) se += expf(logits[i * N + k] - mx);
    const float lse = mx + logf(se);
    // labels index shared memory below: out-of-range values (rejected on the host for host batches; torch's
    // cross_entropy raises) are clamped so that a bad device-resident label cannot read outside the row
    const int yi = min(max((int)y[i], 0), N - 1);
    float pd = 0.f;
    for (int k = 0; k < N; ++k) {
      const float p = expf(logits[i * N + k] - mx) / se;
      prob[i * N + k] = p;
      dl[i * N + k] = (a.mode == HEAD_EXTERNAL_BWD)
                          ? a.dl_ext[(long long)task * a.dl_ext_stride + (long long)(row0 + i) * N + k]
                          : (p - (k == yi ? 1.f : 0.f)) * (wscale * inv_n);
      if (tan) pd += p * ldot[i * N + k];
    }
    if (tan)
      for (int k = 0; k < N; ++k) {
        const float p = prob[i * N + k];
        dldot[i * N + k] = (p * ldot[i * N + k] - p * pd) * inv_n;
      }
    s_rowloss[i] = lse - logits[i * N + yi];
    s_rowcorrect[i] = (am == yi) ? 1.f : 0.f;
  }
  __syncthreads();

  if (a.mode == HEAD_TARGET_FWD) {
    if (tid == 0) {
      float ls = 0.f, cs = 0.f;
      for (int i = 0; i < nl; ++i) { ls += s_rowloss[i]; cs += s_rowcorrect[i]; }
      atomicAdd(&a.loss_out[(long long)task * a.loss_stride], ls * inv_n);       // zeroed at iteration start
      if (a.correct_out) atomicAdd(&a.correct_out[(long long)task * a.correct_stride], cs);
    }
    if (a.logits_out) {
      float* lo = a.logits_out + (long long)task * a.logits_stride + (long long)row0 * N;
      for (int o = tid; o < nl * N; o += 256) lo[o] = logits[o];
    }
    return;
  }

  float* gW = a.gW + (long long)task * a.g_stride + (long long)group * a.g_chunk_stride;
  float* gb = a.gb + (long long)task * a.g_stride + (long long)group * a.g_chunk_stride;
  for (int o = tid; o < N * D; o += 256) {
    const int k = o / D, d = o - k * D;
    float s = 0.f;
    if (!tan) {
      for (int i = 0; i < nl; ++i) s = fmaf(dl[i * N + k], f[(long long)i * D + d], s);
    } else {
      for (int i = 0; i < nl; ++i)
        s += dldot[i * N + k] * f[(long long)i * D + d] + dl[i * N + k] * fd[(long long)i * D + d];
    }
    gW[o] = s;
  }
  if (tid < N) {
    float s = 0.f;
    const float* src = tan ? dldot : dl;
    for (int i = 0; i < nl; ++i) s += src[i * N + tid];
    gb[tid] = s;
  }
  float* df = a.df + (long long)task * a.df_stride + (long long)row0 * D;
  for (int o = tid; o < nl * D; o += 256) {
    const int i = o / D, d = o - i * D;
    float s = 0.f;
    if (!tan) {
      for (int k = 0; k < N; ++k) s = fmaf(dl[i * N + k], W[(long long)k * D + d], s);
    } else {
      for (int k = 0; k < N; ++k)
        s += dldot[i * N + k] * W[(long long)k * D + d] + dl[i * N + k] * uW[(long long)k * D + d];
    }
    df[o] = s;
  }
}

