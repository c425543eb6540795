// Gating maskers: channel (GAP -> MLP -> 2-way decision -> compact index list)
// and spatial/layer (adaptive pool -> 1x1 conv -> decision), plus the mask
// geometry helpers (nearest resize, ExpandMask dilation, row compaction).
//
// Reference behaviour restated (not ported): imagenet_classification/models/
// utils.py:47-65 (Masker_spatial), :74-89 (ExpandMask), :113-131
// (Masker_channel_MLP).  All sums that feed a decision are fp32, fixed order.
#include "laud_common.cuh"

namespace laud {

// ---------------------------------------------------------------------------
// Phase 1 of the deterministic GAP: grid (SPLITS, B), 256 threads.
// x [B,HW,ldx] fp16, C % 8 == 0.  partial[b][s][c] = sum over the rows of split s.
// Threads are laid out as (row lane, 16-byte channel vector) so a warp reads
// contiguous 16B vectors of one pixel row: fully coalesced.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gap_partial_kernel(const __half* __restrict__ x, int HW, int C,
                                                          int ldx, float* __restrict__ partial) {
  __shared__ float red[2048 + 64];
  const int s = blockIdx.x, b = blockIdx.y, nsplit = gridDim.x;
  const int nvec = C >> 3;
  const int vt = nvec < 256 ? nvec : 256;       // vector lanes in use
  const int rl = 256 / vt;                      // row lanes
  const int tid = threadIdx.x;
  const int v0 = tid % vt, r0 = tid / vt;
  const int rows_per = (HW + nsplit - 1) / nsplit;
  const int rbeg = s * rows_per;
  const int rend = min(HW, rbeg + rows_per);
  const __half* xb = x + (size_t)b * HW * ldx;
  float* out = partial + ((size_t)b * nsplit + s) * C;

  for (int vbase = 0; vbase < nvec; vbase += vt) {
    const int v = vbase + v0;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (r0 < rl && v < nvec) {
      for (int r = rbeg + r0; r < rend; r += rl) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)r * ldx + v * 8));
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(h[i]);
          acc[2 * i] += f.x;
          acc[2 * i + 1] += f.y;
        }
      }
    }
    __syncthreads();
    if (r0 < rl && v < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[r0 * (vt * 8) + v0 * 8 + i] = acc[i];
    }
    __syncthreads();
    // fixed-order reduction over the row lanes
    for (int c = tid; c < vt * 8; c += 256) {
      if (vbase * 8 + c < C) {
        float t = 0.f;
        for (int r = 0; r < rl; ++r) t += red[r * (vt * 8) + c];
        out[vbase * 8 + c] = t;
      }
    }
  }
}

// Phase 2 alone (used by laud_global_avg_pool): pooled[b][c] = sum_s partial / HW.
__global__ void gap_final_kernel(const float* __restrict__ partial, int C, int nsplit, int HW,
                                 float* __restrict__ pooled) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int s = 0; s < nsplit; ++s) t += partial[((size_t)b * nsplit + s) * C + c];
  pooled[(size_t)b * C + c] = t / (float)HW;
}

// ---------------------------------------------------------------------------
// Phase 2 + MLP + decision + ordered compaction.  One CTA (256 thr) per sample.
// dynamic smem: p[C] | h[hidden] | l[2G] | flag[G] (as int) | wsum[64]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) masker_decide_kernel(
    const float* __restrict__ partial, const float* __restrict__ pooled_in, int C, int nsplit, int HW,
    int layers, const float* __restrict__ w1, const float* __restrict__ b1, int hidden,
    const float* __restrict__ w2, const float* __restrict__ b2, int G,
    float* __restrict__ pooled_out, float* __restrict__ logits_out, uint8_t* __restrict__ mask_out,
    int* __restrict__ idx_out, int* __restrict__ cnt_out, int* __restrict__ total_out) {
  extern __shared__ float sm[];
  float* p = sm;
  float* h = p + C;
  float* l = h + (hidden > 0 ? hidden : 1);
  int* flag = reinterpret_cast<int*>(l + 2 * G);
  int* wsum = flag + G;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int c = tid; c < C; c += 256) {
    float t;
    if (pooled_in) {
      t = pooled_in[(size_t)b * C + c];
    } else {
      t = 0.f;
      for (int s = 0; s < nsplit; ++s) t += partial[((size_t)b * nsplit + s) * C + c];
      t = t / (float)HW;
    }
    p[c] = t;
    if (pooled_out) pooled_out[(size_t)b * C + c] = t;
  }
  __syncthreads();

  // first linear layer: one warp per output row, lanes stride the channels
  const int rows1 = layers == 2 ? hidden : 2 * G;
  float* dst1 = layers == 2 ? h : l;
  for (int j = warp; j < rows1; j += 8) {
    const float* wr = w1 + (size_t)j * C;
    float t = 0.f;
    for (int c = lane; c < C; c += 32) t = fmaf(wr[c], p[c], t);
    t = warp_sum(t);
    if (lane == 0) {
      t += b1[j];
      dst1[j] = layers == 2 ? fmaxf(t, 0.f) : t;
    }
  }
  __syncthreads();
  if (layers == 2) {
    for (int o = tid; o < 2 * G; o += 256) {
      const float* wr = w2 + (size_t)o * hidden;
      float t = 0.f;
      for (int j = 0; j < hidden; ++j) t = fmaf(wr[j], h[j], t);
      l[o] = t + b2[o];
    }
    __syncthreads();
  }
  for (int o = tid; o < 2 * G; o += 256)
    if (logits_out) logits_out[(size_t)b * 2 * G + o] = l[o];
  for (int g = tid; g < G; g += 256) {
    const int f = l[g] >= l[G + g] ? 1 : 0;          // ties keep (utils.py:127)
    flag[g] = f;
    mask_out[(size_t)b * G + g] = (uint8_t)f;
  }
  __syncthreads();

  // ordered compaction: chunks of 256 groups; ballot + warp prefix
  int base_on = 0;
  // pass 1: active ids
  for (int g0 = 0; g0 < G; g0 += 256) {
    const int g = g0 + tid;
    const int f = g < G ? flag[g] : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base_on;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (f) idx_out[(size_t)b * G + off + __popc(bal & ((1u << lane) - 1))] = g;
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += wsum[w];
    base_on += tot;
    __syncthreads();
  }
  // pass 2: inactive ids behind them
  int base_off = base_on;
  for (int g0 = 0; g0 < G; g0 += 256) {
    const int g = g0 + tid;
    const int f = g < G ? (flag[g] ? 0 : 1) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base_off;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (f) idx_out[(size_t)b * G + off + __popc(bal & ((1u << lane) - 1))] = g;
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += wsum[w];
    base_off += tot;
    __syncthreads();
  }
  if (tid == 0) {
    cnt_out[b] = base_on;
    if (total_out) atomicAdd(total_out, base_on);
  }
}

// ---------------------------------------------------------------------------
// Fused channel masker (the product path of laud_masker_channel_mlp): ONE launch, one CTA of 512
// threads per sample: GAP (16-byte loads, 4 rows in flight per thread, fixed-order fp32 reduction)
// -> MLP -> keep>=drop -> ordered compaction.  Replaces the gap_partial + masker_decide pair (two
// launches, a [B,8,C] round trip through HBM and ~60 us of exposed latency per block).
// dynamic smem: red[512*8] | p[C] | h[hidden] | l[2G] | flag[G] | wsum[64]
// ---------------------------------------------------------------------------
constexpr int MF_THREADS = 512;
__global__ void __launch_bounds__(MF_THREADS) masker_channel_fused_kernel(
    const __half* __restrict__ x, int HW, int C, int layers, const float* __restrict__ w1,
    const float* __restrict__ b1, int hidden, const float* __restrict__ w2, const float* __restrict__ b2, int G,
    float* __restrict__ pooled_out, float* __restrict__ logits_out, uint8_t* __restrict__ mask_out,
    int* __restrict__ idx_out, int* __restrict__ cnt_out, int* __restrict__ total_out,
    const float* __restrict__ part, int gap_tiles) {
  extern __shared__ float sm[];
  float* red = sm;
  float* p = red + MF_THREADS * 8;
  float* h = p + C;
  float* l = h + (hidden > 0 ? hidden : 1);
  int* flag = reinterpret_cast<int*>(l + 2 * G);
  int* wsum = flag + G;
  constexpr int NW = MF_THREADS / 32;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvec = C >> 3;
  const int vt = nvec < MF_THREADS ? nvec : MF_THREADS;      // vector lanes in use
  const int rl = MF_THREADS / vt;                            // row lanes
  const int v0 = tid % vt, r0 = tid / vt;
  const __half* xb = x + (size_t)b * HW * C;
  const float inv = 1.0f / (float)HW;
  // programmatic dependent launch: our successor may start its prologue; we wait for our predecessor's writes
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (part) {
    // pooled features from the partial sums the producing convolution left (laud_conv_desc::gap_partial):
    // the tiles of 128 consecutive pixels of the flat [B*HW] list that hold pixels of sample b, ascending
    const int t_first = (int)(((long long)b * HW) / 128), t_last = (int)(((long long)(b + 1) * HW - 1) / 128);
    const int nk = min(t_last - t_first + 1, gap_tiles);
    const float* pb = part + (size_t)b * gap_tiles * C;
    for (int c = tid; c < C; c += MF_THREADS) {
      float t = 0.f;
      if (nk <= 4) {                                         // all partials of the channel in flight at once; same ascending order
        float q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = k < nk ? __ldg(pb + (size_t)k * C + c) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nk) t += q[k];
      } else {
#pragma unroll 4
        for (int k = 0; k < nk; ++k) t += __ldg(pb + (size_t)k * C + c);
      }
      t *= inv;
      p[c] = t;
      if (pooled_out) pooled_out[(size_t)b * C + c] = t;
    }
  } else
  for (int vbase = 0; vbase < nvec; vbase += vt) {
    const int v = vbase + v0;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (r0 < rl && v < nvec) {
      const uint4* src = reinterpret_cast<const uint4*>(xb + v * 8);
      const size_t pitch = (size_t)C >> 3;                   // row pitch in 16-byte units
      int r = r0;
      for (; r + 3 * rl < HW; r += 4 * rl) {                 // four independent loads in flight
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = __ldg(src + (size_t)(r + u * rl) * pitch);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const __half2* hh = reinterpret_cast<const __half2*>(&q[u]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(hh[i]);
            acc[2 * i] += f.x;
            acc[2 * i + 1] += f.y;
          }
        }
      }
      for (; r < HW; r += rl) {
        const uint4 q = __ldg(src + (size_t)r * pitch);
        const __half2* hh = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(hh[i]);
          acc[2 * i] += f.x;
          acc[2 * i + 1] += f.y;
        }
      }
    }
    __syncthreads();
    if (r0 < rl && v < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[r0 * (vt * 8) + v0 * 8 + i] = acc[i];
    }
    __syncthreads();
    for (int c = tid; c < vt * 8; c += MF_THREADS) {         // fixed-order reduction over the row lanes
      if (vbase * 8 + c < C) {
        float t = 0.f;
        for (int r = 0; r < rl; ++r) t += red[r * (vt * 8) + c];
        t *= inv;
        p[vbase * 8 + c] = t;
        if (pooled_out) pooled_out[(size_t)b * C + vbase * 8 + c] = t;
      }
    }
  }
  __syncthreads();

  // first linear layer: one warp per output row, lanes stride the channels
  const int rows1 = layers == 2 ? hidden : 2 * G;
  float* dst1 = layers == 2 ? h : l;
  for (int j = warp; j < rows1; j += NW) {
    const float* wr = w1 + (size_t)j * C;
    float t = 0.f;
    int c = lane;
    for (; c + 224 < C; c += 256) {                          // eight loads in flight; the fmaf chain keeps its order
      float wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = __ldg(wr + c + 32 * u);
#pragma unroll
      for (int u = 0; u < 8; ++u) t = fmaf(wv[u], p[c + 32 * u], t);
    }
    for (; c < C; c += 32) t = fmaf(__ldg(wr + c), p[c], t);
    t = warp_sum(t);
    if (lane == 0) {
      t += b1[j];
      dst1[j] = layers == 2 ? fmaxf(t, 0.f) : t;
    }
  }
  __syncthreads();
  if (layers == 2) {
    for (int o = tid; o < 2 * G; o += MF_THREADS) {
      const float* wr = w2 + (size_t)o * hidden;
      float t = 0.f;
      if ((hidden & 3) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0) {
        // whole row in flight (hidden is 16 for every LAUD-ResNet stage); the fmaf chain keeps its order
        const float4* w4 = reinterpret_cast<const float4*>(wr);
        int j = 0;
        for (; j + 16 <= hidden; j += 16) {
          const float4 q0 = __ldg(w4 + (j >> 2)), q1 = __ldg(w4 + (j >> 2) + 1), q2 = __ldg(w4 + (j >> 2) + 2), q3 = __ldg(w4 + (j >> 2) + 3);
          t = fmaf(q0.x, h[j], t); t = fmaf(q0.y, h[j + 1], t); t = fmaf(q0.z, h[j + 2], t); t = fmaf(q0.w, h[j + 3], t);
          t = fmaf(q1.x, h[j + 4], t); t = fmaf(q1.y, h[j + 5], t); t = fmaf(q1.z, h[j + 6], t); t = fmaf(q1.w, h[j + 7], t);
          t = fmaf(q2.x, h[j + 8], t); t = fmaf(q2.y, h[j + 9], t); t = fmaf(q2.z, h[j + 10], t); t = fmaf(q2.w, h[j + 11], t);
          t = fmaf(q3.x, h[j + 12], t); t = fmaf(q3.y, h[j + 13], t); t = fmaf(q3.z, h[j + 14], t); t = fmaf(q3.w, h[j + 15], t);
        }
        for (; j < hidden; j += 4) {
          const float4 q0 = __ldg(w4 + (j >> 2));
          t = fmaf(q0.x, h[j], t); t = fmaf(q0.y, h[j + 1], t); t = fmaf(q0.z, h[j + 2], t); t = fmaf(q0.w, h[j + 3], t);
        }
      } else {
        for (int j = 0; j < hidden; ++j) t = fmaf(__ldg(wr + j), h[j], t);
      }
      l[o] = t + b2[o];
    }
    __syncthreads();
  }
  for (int o = tid; o < 2 * G; o += MF_THREADS)
    if (logits_out) logits_out[(size_t)b * 2 * G + o] = l[o];
  for (int g = tid; g < G; g += MF_THREADS) {
    const int f = l[g] >= l[G + g] ? 1 : 0;          // ties keep (utils.py:127)
    flag[g] = f;
    mask_out[(size_t)b * G + g] = (uint8_t)f;
  }
  __syncthreads();

  // ordered compaction (active ids, then inactive ids): ballot + warp prefix over chunks of 512 groups
  int base = 0, n_on = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int g0 = 0; g0 < G; g0 += MF_THREADS) {
      const int g = g0 + tid;
      const int f = g < G ? (flag[g] ? 1 - pass : pass) : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, f);
      if (lane == 0) wsum[warp] = __popc(bal);
      __syncthreads();
      int off = base, tot = 0;
      for (int w = 0; w < NW; ++w) {
        if (w < warp) off += wsum[w];
        tot += wsum[w];
      }
      if (f) idx_out[(size_t)b * G + off + __popc(bal & ((1u << lane) - 1))] = g;
      base += tot;
      __syncthreads();
    }
    if (pass == 0) n_on = base;
  }
  if (tid == 0) {
    cnt_out[b] = n_on;
    if (total_out) atomicAdd(total_out, n_on);
  }
}

// ---------------------------------------------------------------------------
// Spatial masker: one warp per (b, cell).  Pool the cell (adaptive_avg_pool2d
// region: [floor(i*H/S), ceil((i+1)*H/S)) ), then the 2g-row 1x1 conv.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) masker_spatial_kernel(
    const __half* __restrict__ x, int B, int H, int W, int C, const float* __restrict__ w,
    const float* __restrict__ bias, int g, int S, float* __restrict__ logits_out,
    uint8_t* __restrict__ mask_out, int* __restrict__ total_out) {
  // lanes = (pixel lane, 16-byte channel vector): narrow layers (C = 32) still issue full-width loads
  __shared__ int s_ones;
  if (threadIdx.x == 0) s_ones = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long cell = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool live = cell < (long long)B * S * S;
  int ones = 0;
  if (live) {
    const int b = (int)(cell / (S * S));
    const int ci = (int)(cell % (S * S)) / S, cj = (int)(cell % S);
    int y0, y1, x0, x1;
    if (S >= H) { y0 = ci; y1 = ci + 1; x0 = cj; x1 = cj + 1; }
    else {
      y0 = (ci * H) / S; y1 = ((ci + 1) * H + S - 1) / S;
      x0 = (cj * W) / S; x1 = ((cj + 1) * W + S - 1) / S;
    }
    const int wreg = x1 - x0, npix = (y1 - y0) * wreg;
    const float area = (float)npix;
    int nv = 1;
    while (nv < 32 && nv * 8 < C) nv <<= 1;             // vector lanes per pixel (power of two)
    const int pl = 32 / nv, vl = lane % nv, pi = lane / nv;
    float part[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) part[i] = 0.f;
    const int G2 = 2 * g;
    for (int cb = 0; cb < C; cb += nv * 8) {
      const int c0 = cb + vl * 8;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      if (c0 < C)
        for (int idx = pi; idx < npix; idx += pl) {
          const int yy = y0 + idx / wreg, xx = x0 + idx % wreg;
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + yy) * W + xx) * C + c0));
          const __half2* hh = reinterpret_cast<const __half2*>(&q);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(hh[i]);
            acc[2 * i] += f.x;
            acc[2 * i + 1] += f.y;
          }
        }
      for (int off = nv; off < 32; off <<= 1) {          // fixed-order tree over the pixel lanes
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
      }
      if (pi == 0 && c0 < C) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = (S >= H) ? acc[i] : acc[i] / area;
        for (int o = 0; o < G2 && o < 8; ++o) {
          const float* wr = w + (size_t)o * C + c0;
          float t = part[o];
#pragma unroll
          for (int i = 0; i < 8; ++i) t = fmaf(wr[i], acc[i], t);
          part[o] = t;
        }
      }
    }
    for (int o = 0; o < G2 && o < 8; ++o) part[o] = warp_sum(part[o]) + bias[o];
    if (lane == 0) {
      for (int o = 0; o < G2; ++o)
        if (logits_out) logits_out[(((size_t)b * G2 + o) * S + ci) * S + cj] = part[o];
      for (int q = 0; q < g; ++q) {
        const int f = part[q] >= part[g + q] ? 1 : 0;
        mask_out[(((size_t)b * g + q) * S + ci) * S + cj] = (uint8_t)f;
        ones += f;
      }
      if (ones) atomicAdd(&s_ones, ones);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && total_out && s_ones) atomicAdd(total_out, s_ones);   // one global atomic per CTA
}

__global__ void resize_mask_kernel(const uint8_t* __restrict__ m, int B, int g, int S, int Ho,
                                   uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * g * Ho * Ho;
  if (i >= n) return;
  const int ox = (int)(i % Ho), oy = (int)((i / Ho) % Ho);
  const long long bg = i / ((long long)Ho * Ho);
  const int sy = (oy * S) / Ho, sx = (ox * S) / Ho;       // floor(dst * in / out)
  out[i] = m[(bg * S + sy) * S + sx];
}

// ExpandMask: out[b,*,y,x] = OR over groups and window of zero-inserted mask.
__global__ void expand_mask_kernel(const uint8_t* __restrict__ m, int B, int g, int H, int W,
                                   int stride, int pad, uint8_t* __restrict__ out,
                                   int* __restrict__ total_out) {
  const int Ho = H * stride, Wo = W * stride;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * Ho * Wo;
  int f = 0;
  if (i < n) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const int b = (int)(i / ((long long)Ho * Wo));
    for (int dy = -pad; dy <= pad && !f; ++dy)
      for (int dx = -pad; dx <= pad && !f; ++dx) {
        const int uy = oy + dy, ux = ox + dx;
        if (uy < 0 || ux < 0 || uy >= Ho || ux >= Wo) continue;
        if (uy % stride || ux % stride) continue;      // zero-inserted positions
        for (int q = 0; q < g; ++q)
          if (m[(((size_t)b * g + q) * H + uy / stride) * W + ux / stride]) { f = 1; break; }
      }
    for (int q = 0; q < g; ++q) out[(((size_t)b * g + q) * Ho + oy) * Wo + ox] = (uint8_t)f;
  }
  // count ones (x g groups, as the reference's mean runs over [B,g,H,W])
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && bal && total_out) atomicAdd(total_out, __popc(bal) * g);
}

// The three spatial masks of a block in ONE launch (laud_resnet.py:105-110): m3 = nearest-resize(small) to the output
// size, m2 = ExpandMask(1, 0)(m3) (the OR over mask groups), m1 = ExpandMask(stride, 1)(m2) at the input size, with the
// counts the statistics need.  Everything is a function of the tiny gate map `small`, so each thread derives its cells
// from it directly.  Thread i < B*Hi*Hi owns input-resolution cell i (m1) and, if i < B*Ho*Ho, output cell i (m3, m2).
__device__ __forceinline__ int m2_at(const uint8_t* __restrict__ small, int b, int g, int S, int Ho, int y, int x) {
  const int sy = (y * S) / Ho, sx = (x * S) / Ho;
  for (int q = 0; q < g; ++q)
    if (small[(((size_t)b * g + q) * S + sy) * S + sx]) return 1;
  return 0;
}
__global__ void spatial_masks_kernel(const uint8_t* __restrict__ small, int B, int g, int S, int Ho, int stride,
                                     uint8_t* __restrict__ m3, uint8_t* __restrict__ m2, uint8_t* __restrict__ m1,
                                     int* __restrict__ total2, int* __restrict__ total1) {
  const int Hi = Ho * stride;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_out = (long long)B * Ho * Ho, n_in = (long long)B * Hi * Hi;
  int f2 = 0, f1 = 0;
  if (i < n_out) {
    const int x = (int)(i % Ho), y = (int)((i / Ho) % Ho), b = (int)(i / ((long long)Ho * Ho));
    const int sy = (y * S) / Ho, sx = (x * S) / Ho;
    for (int q = 0; q < g; ++q) {
      const uint8_t v = small[(((size_t)b * g + q) * S + sy) * S + sx];
      m3[(((size_t)b * g + q) * Ho + y) * Ho + x] = v;
      f2 |= v != 0;
    }
    for (int q = 0; q < g; ++q) m2[(((size_t)b * g + q) * Ho + y) * Ho + x] = (uint8_t)f2;
  }
  if (i < n_in) {
    const int X = (int)(i % Hi), Y = (int)((i / Hi) % Hi), b = (int)(i / ((long long)Hi * Hi));
    for (int dy = -1; dy <= 1 && !f1; ++dy)
      for (int dx = -1; dx <= 1 && !f1; ++dx) {
        const int uy = Y + dy, ux = X + dx;
        if (uy < 0 || ux < 0 || uy >= Hi || ux >= Hi || uy % stride || ux % stride) continue;
        f1 = m2_at(small, b, g, S, Ho, uy / stride, ux / stride);
      }
    for (int q = 0; q < g; ++q) m1[(((size_t)b * g + q) * Hi + Y) * Hi + X] = (uint8_t)f1;
  }
  const unsigned b2 = __ballot_sync(0xffffffffu, f2), b1 = __ballot_sync(0xffffffffu, f1);
  if ((threadIdx.x & 31) == 0) {
    if (b2 && total2) atomicAdd(total2, __popc(b2) * g);
    if (b1 && total1) atomicAdd(total1, __popc(b1) * g);
  }
}

extern "C" int laud_spatial_masks(const uint8_t* small, int B, int g, int S, int H_out, int stride, uint8_t* m3, uint8_t* m2,
                                  uint8_t* m1, int32_t* total2, int32_t* total1, void* stream) {
  LAUD_REQUIRE(small && m3 && m2 && m1 && B > 0 && g > 0 && S > 0 && H_out > 0 && stride >= 1,
               "laud_spatial_masks: bad arguments");
  const long long n = (long long)B * H_out * stride * H_out * stride;
  spatial_masks_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(small, B, g, S, H_out, stride, m3, m2, m1,
                                                                                 total2, total1);
  return check_launch("spatial_masks_kernel");
}

// Ordered compaction of active rows.  Pass A: per-CTA counts (2048 items each);
// pass B: one CTA scans the counts; pass C: ordered scatter.
constexpr int kCompactItems = 2048;
__device__ __forceinline__ int row_active(const uint8_t* gate, int g, int HW, long long i) {
  if (g == 1) return gate[i] != 0;
  const long long b = i / HW, p = i % HW;
  for (int q = 0; q < g; ++q)
    if (gate[(b * g + q) * HW + p]) return 1;
  return 0;
}
__global__ void __launch_bounds__(256) compact_count_kernel(const uint8_t* gate, int g, int HW, long long n,
                                                            int* block_ws) {
  __shared__ int ws[8];
  int c = 0;
  const long long base = (long long)blockIdx.x * kCompactItems;
  for (int k = 0; k < kCompactItems / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    c += (i < n) ? row_active(gate, g, HW, i) : 0;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    block_ws[blockIdx.x] = t;
  }
}
__global__ void compact_scan_kernel(int* block_ws, int nblk, int* count_out) {
  // single thread: nblk is small (<= B*HW/2048)
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < nblk; ++i) { const int t = block_ws[i]; block_ws[i] = run; run += t; }
    count_out[0] = run;
  }
}
__global__ void __launch_bounds__(256) compact_write_kernel(const uint8_t* gate, int g, int HW, long long n,
                                                            const int* block_ws, int* rows_out) {
  __shared__ int ws[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int run = block_ws[blockIdx.x];
  const long long base = (long long)blockIdx.x * kCompactItems;
  for (int k = 0; k < kCompactItems / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    const int f = (i < n) ? row_active(gate, g, HW, i) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) ws[warp] = __popc(bal);
    __syncthreads();
    int off = run;
    for (int w = 0; w < warp; ++w) off += ws[w];
    if (f) rows_out[off + __popc(bal & ((1u << lane) - 1))] = (int)i;
    for (int w = 0; w < 8; ++w) run += ws[w];
    __syncthreads();
  }
}

}  // namespace laud

using namespace laud;

static size_t decide_smem(int C, int hidden, int G) {
  return sizeof(float) * ((size_t)C + (hidden > 0 ? hidden : 1) + 2 * (size_t)G) + sizeof(int) * ((size_t)G + 64);
}

extern "C" int laud_global_avg_pool(const void* x, int B, int HW, int C, int ldx, float* partial_ws,
                                    float* pooled_out, void* stream) {
  LAUD_REQUIRE(x && partial_ws && pooled_out, "laud_global_avg_pool: null pointer");
  LAUD_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldx >= C,
               "laud_global_avg_pool: need C %% 8 == 0 and ldx %% 8 == 0 (C=%d ldx=%d)", C, ldx);
  cudaStream_t s = (cudaStream_t)stream;
  gap_partial_kernel<<<dim3(LAUD_GAP_SPLITS, B), 256, 0, s>>>((const __half*)x, HW, C, ldx, partial_ws);
  if (int e = check_launch("gap_partial_kernel")) return e;
  gap_final_kernel<<<dim3((C + 255) / 256, B), 256, 0, s>>>(partial_ws, C, LAUD_GAP_SPLITS, HW, pooled_out);
  return check_launch("gap_final_kernel");
}

static int launch_decide(const float* partial, const float* pooled_in, int B, int HW, int C, int layers,
                         const float* w1, const float* b1, int hidden, const float* w2, const float* b2,
                         int G, float* pooled_out, float* logits_out, uint8_t* mask_out, int32_t* idx_out,
                         int32_t* cnt_out, int32_t* total_out, cudaStream_t s) {
  LAUD_REQUIRE(layers == 1 || layers == 2, "channel masker: layers must be 1 or 2 (got %d)", layers);
  LAUD_REQUIRE(w1 && b1 && mask_out && idx_out && cnt_out, "channel masker: null pointer");
  LAUD_REQUIRE(layers == 1 || (w2 && b2 && hidden > 0), "channel masker: 2-layer MLP needs w2,b2,hidden");
  LAUD_REQUIRE(G > 0, "channel masker: G must be positive");
  const size_t smem = decide_smem(C, layers == 2 ? hidden : 0, G);
  LAUD_REQUIRE(smem <= 200 * 1024, "channel masker: C/G too large for shared memory");
  static size_t decide_smem_set[MAX_DEVICES] = {0};    // per device; sizes up to 48 KB need no attribute
  const int dev = current_device();
  if (smem > 48 * 1024 && smem > decide_smem_set[dev]) {
    LAUD_CUDA(cudaFuncSetAttribute(masker_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decide_smem_set[dev] = smem;
  }
  masker_decide_kernel<<<B, 256, smem, s>>>(partial, pooled_in, C, LAUD_GAP_SPLITS, HW, layers, w1, b1,
                                            layers == 2 ? hidden : 0, w2, b2, G, pooled_out, logits_out,
                                            mask_out, idx_out, cnt_out, total_out);
  return check_launch("masker_decide_kernel");
}

static size_t g_fused_smem_set[MAX_DEVICES] = {0};     // largest dynamic shared memory size set on the fused kernel so far, per device

// launched with programmatic stream serialization: the kernel's first instructions overlap the predecessor's tail
static void launch_masker_fused(int B, size_t smem, cudaStream_t s, const __half* x, int HW, int C, int layers,
                                const float* w1, const float* b1, int hidden, const float* w2, const float* b2, int G,
                                float* pooled_out, float* logits_out, uint8_t* mask_out, int* idx_out, int* cnt_out,
                                int* total_out, const float* part, int gap_tiles) {
  static const bool pdl = getenv("LAUD_PDL") != nullptr;   // opt-in: measured -2.5 % with the two graph chains (early CTAs of one chain sit on SMs the other chain could use), +0.8 % with one
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)B);
  cfg.blockDim = dim3(MF_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, masker_channel_fused_kernel, x, HW, C, layers, w1, b1, hidden, w2, b2, G, pooled_out, logits_out,
                     mask_out, idx_out, cnt_out, total_out, part, gap_tiles);
}

extern "C" int laud_masker_channel_mlp(const void* x, int B, int HW, int C, int layers, const float* w1,
                                       const float* b1, int hidden, const float* w2, const float* b2, int G,
                                       float* partial_ws, float* pooled_out, float* logits_out,
                                       uint8_t* mask_out, int32_t* idx_out, int32_t* cnt_out,
                                       int32_t* total_out, void* stream) {
  LAUD_REQUIRE(x && partial_ws, "laud_masker_channel_mlp: null pointer");
  LAUD_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0, "laud_masker_channel_mlp: need C %% 8 == 0 (C=%d)", C);
  cudaStream_t s = (cudaStream_t)stream;
  LAUD_REQUIRE(layers == 1 || layers == 2, "channel masker: layers must be 1 or 2 (got %d)", layers);
  LAUD_REQUIRE(w1 && b1 && mask_out && idx_out && cnt_out, "channel masker: null pointer");
  LAUD_REQUIRE(layers == 1 || (w2 && b2 && hidden > 0), "channel masker: 2-layer MLP needs w2,b2,hidden");
  LAUD_REQUIRE(G > 0, "channel masker: G must be positive");
  const size_t smem = sizeof(float) * MF_THREADS * 8 + decide_smem(C, layers == 2 ? hidden : 0, G);
  LAUD_REQUIRE(smem <= 200 * 1024, "channel masker: C/G too large for shared memory");
  if (smem > 48 * 1024 && smem > g_fused_smem_set[current_device()]) {
    LAUD_CUDA(cudaFuncSetAttribute(masker_channel_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    g_fused_smem_set[current_device()] = smem;
  }
  launch_masker_fused(B, smem, s, (const __half*)x, HW, C, layers, w1, b1, layers == 2 ? hidden : 0, w2, b2, G, pooled_out,
                      logits_out, mask_out, idx_out, cnt_out, total_out, nullptr, 0);
  return check_launch("masker_channel_fused_kernel");
}

extern "C" int laud_masker_channel_from_partials(const float* partials, int B, int HW, int C, int gap_tiles, int layers,
                                                 const float* w1, const float* b1, int hidden, const float* w2,
                                                 const float* b2, int G, float* pooled_out, float* logits_out,
                                                 uint8_t* mask_out, int32_t* idx_out, int32_t* cnt_out,
                                                 int32_t* total_out, void* stream) {
  LAUD_REQUIRE(partials, "laud_masker_channel_from_partials: null pointer");
  LAUD_REQUIRE(B > 0 && HW > 0 && C > 0 && C % 8 == 0, "laud_masker_channel_from_partials: need C %% 8 == 0 (C=%d)", C);
  LAUD_REQUIRE(gap_tiles >= (HW - 1) / 128 + 2, "laud_masker_channel_from_partials: gap_tiles %d < (HW-1)/128 + 2", gap_tiles);
  cudaStream_t s = (cudaStream_t)stream;
  LAUD_REQUIRE(layers == 1 || layers == 2, "channel masker: layers must be 1 or 2 (got %d)", layers);
  LAUD_REQUIRE(w1 && b1 && mask_out && idx_out && cnt_out, "channel masker: null pointer");
  LAUD_REQUIRE(layers == 1 || (w2 && b2 && hidden > 0), "channel masker: 2-layer MLP needs w2,b2,hidden");
  LAUD_REQUIRE(G > 0, "channel masker: G must be positive");
  const size_t smem = sizeof(float) * MF_THREADS * 8 + decide_smem(C, layers == 2 ? hidden : 0, G);
  LAUD_REQUIRE(smem <= 200 * 1024, "channel masker: C/G too large for shared memory");
  if (smem > 48 * 1024 && smem > g_fused_smem_set[current_device()]) {
    LAUD_CUDA(cudaFuncSetAttribute(masker_channel_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    g_fused_smem_set[current_device()] = smem;
  }
  launch_masker_fused(B, smem, s, nullptr, HW, C, layers, w1, b1, layers == 2 ? hidden : 0, w2, b2, G, pooled_out, logits_out,
                      mask_out, idx_out, cnt_out, total_out, partials, gap_tiles);
  return check_launch("masker_channel_fused_kernel");
}

extern "C" int laud_masker_channel_from_pooled(const float* pooled, int B, int C, int layers, const float* w1,
                                               const float* b1, int hidden, const float* w2, const float* b2,
                                               int G, float* logits_out, uint8_t* mask_out, int32_t* idx_out,
                                               int32_t* cnt_out, int32_t* total_out, void* stream) {
  LAUD_REQUIRE(pooled && B > 0 && C > 0, "laud_masker_channel_from_pooled: bad arguments");
  return launch_decide(nullptr, pooled, B, 1, C, layers, w1, b1, hidden, w2, b2, G, nullptr, logits_out,
                       mask_out, idx_out, cnt_out, total_out, (cudaStream_t)stream);
}

extern "C" int laud_masker_spatial(const void* x, int B, int H, int W, int C, const float* w, const float* bias,
                                   int g, int S, float* logits_out, uint8_t* mask_out, int32_t* total_out,
                                   void* stream) {
  LAUD_REQUIRE(x && w && bias && mask_out, "laud_masker_spatial: null pointer");
  LAUD_REQUIRE(B > 0 && H > 0 && W > 0 && C % 8 == 0 && S > 0, "laud_masker_spatial: bad shape (C=%d S=%d)", C, S);
  LAUD_REQUIRE(g >= 1 && g <= 4, "laud_masker_spatial: mask_channel_group must be in [1,4] (got %d)", g);
  LAUD_REQUIRE(S <= H, "laud_masker_spatial: mask_size %d exceeds feature size %d", S, H);
  const long long cells = (long long)B * S * S;
  const int grid = (int)((cells + 7) / 8);
  masker_spatial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)x, B, H, W, C, w, bias, g, S,
                                                                logits_out, mask_out, total_out);
  return check_launch("masker_spatial_kernel");
}

extern "C" int laud_resize_mask_nearest(const uint8_t* mask, int B, int g, int S, int H_out, uint8_t* out,
                                        void* stream) {
  LAUD_REQUIRE(mask && out && B > 0 && g > 0 && S > 0 && H_out > 0, "laud_resize_mask_nearest: bad arguments");
  const long long n = (long long)B * g * H_out * H_out;
  resize_mask_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mask, B, g, S, H_out, out);
  return check_launch("resize_mask_kernel");
}

// A per-(sample, group) gate broadcast over hw positions (the S == 1 case of the nearest resize on a NON-square map:
// lad_mmdet_resnet.py:274 resizes the mask to the actual feature size); total_out += ones written (nullable).
namespace laud {
namespace {
__global__ void broadcast_gate_kernel(const uint8_t* __restrict__ gate, long long n, int hw, uint8_t* __restrict__ out,
                                      int* __restrict__ total) {
  int ones = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t v = gate[i / hw] ? 1 : 0;
    out[i] = v;
    ones += v;
  }
  if (total) {
    ones = (int)warp_sum((float)ones);                          // (exact: at most 32 x a few thousand per warp)
    if ((threadIdx.x & 31) == 0 && ones) atomicAdd(total, ones);
  }
}
}  // namespace
}  // namespace laud

extern "C" int laud_broadcast_gate(const uint8_t* gate, int B, int g, int hw, uint8_t* out, int32_t* total_out, void* stream) {
  LAUD_REQUIRE(gate && out && B > 0 && g > 0 && hw > 0, "laud_broadcast_gate: bad arguments");
  const long long n = (long long)B * g * hw;
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  broadcast_gate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gate, n, hw, out, total_out);
  return check_launch("broadcast_gate_kernel");
}

extern "C" int laud_expand_mask(const uint8_t* mask, int B, int g, int H, int W, int stride, int padding,
                                uint8_t* out, int32_t* total_out, void* stream) {
  LAUD_REQUIRE(mask && out && B > 0 && g > 0 && H > 0 && W > 0, "laud_expand_mask: bad arguments");
  LAUD_REQUIRE(stride >= 1 && padding >= 0 && padding <= 3, "laud_expand_mask: bad stride/padding");
  const long long n = (long long)B * H * stride * W * stride;
  expand_mask_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mask, B, g, H, W, stride, padding,
                                                                              out, total_out);
  return check_launch("expand_mask_kernel");
}

// Layer gate bookkeeping in one launch (one CTA): ordered list of the ACTIVE samples of a [B] gate + the counts the
// statistics need (ones of the gate broadcast to the H_out x W_out and, dilated, to the H_in x W_in maps: a per-sample
// gate dilates to itself).  Replaces resize + 2 x expand + 3-kernel compaction for dyn_mode='layer'.
__global__ void __launch_bounds__(1024) layer_gate_lists_kernel(const uint8_t* __restrict__ gate, int B, int hw_out,
                                                                int hw_in, int* __restrict__ counts4,
                                                                int* __restrict__ rows_out, int* __restrict__ count_out) {
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int base = 0;
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    const int f = (b < B && gate[b]) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base, tot = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) off += wsum[w];
      tot += wsum[w];
    }
    if (f) rows_out[off + __popc(bal & ((1u << lane) - 1))] = b;
    base += tot;
    __syncthreads();
  }
  if (tid == 0) {
    *count_out = base;
    if (counts4) {
      counts4[2] += base * hw_out;
      counts4[3] += base * hw_in;
    }
  }
}

extern "C" int laud_layer_gate_lists(const uint8_t* gate, int B, int hw_out, int hw_in, int32_t* counts4,
                                     int32_t* rows_out, int32_t* count_out, void* stream) {
  LAUD_REQUIRE(gate && rows_out && count_out && B > 0, "laud_layer_gate_lists: bad arguments");
  layer_gate_lists_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(gate, B, hw_out, hw_in, counts4, rows_out, count_out);
  return check_launch("layer_gate_lists_kernel");
}

extern "C" int laud_compact_rows(const uint8_t* gate, int B, int g, int HW, int32_t* rows_out,
                                 int32_t* count_out, int32_t* block_ws, void* stream) {
  LAUD_REQUIRE(gate && rows_out && count_out && block_ws && B > 0 && g > 0 && HW > 0,
               "laud_compact_rows: bad arguments");
  const long long n = (long long)B * HW;
  LAUD_REQUIRE(n < (1ll << 31), "laud_compact_rows: too many rows");
  const int nblk = (int)((n + kCompactItems - 1) / kCompactItems);
  cudaStream_t s = (cudaStream_t)stream;
  compact_count_kernel<<<nblk, 256, 0, s>>>(gate, g, HW, n, block_ws);
  if (int e = check_launch("compact_count_kernel")) return e;
  compact_scan_kernel<<<1, 32, 0, s>>>(block_ws, nblk, count_out);
  if (int e = check_launch("compact_scan_kernel")) return e;
  compact_write_kernel<<<nblk, 256, 0, s>>>(gate, g, HW, n, block_ws, rows_out);
  return check_launch("compact_write_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Training-mode gate with SUPPLIED Gumbel noise (reference utils.py:56-58, 123-125, 161-163):
//   F.gumbel_softmax(logits.view(b, 2, ...), dim=1, tau, hard=True)[:, 0]
// draws g ~ Gumbel(0,1) from torch's generator, forms y = softmax((logits + g) / tau) over the keep/drop pair and returns
// the one-hot argmax (index 0 = keep wins ties).  A custom kernel cannot reproduce torch's Philox stream, so the noise is
// an INPUT: given the same noise tensor the forward value of the gate is reproduced exactly:
//   keep  <=>  (l_keep + g_keep) / tau  >=  (l_drop + g_drop) / tau        (fp32, the reference's operation order)
// logits / noise: fp32 [B, 2, G, inner] (channel gates: inner = 1; spatial gates: inner = S*S).
// One CTA per sample; with idx_out the active group ids are compacted in ascending order, then the inactive ones (the
// layout the eval-mode maskers emit), cnt_out[b] = #active.  total_out += #active decisions (statistics counter).
__global__ void __launch_bounds__(256) gate_from_logits_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                                                               int G, int inner, float tau, uint8_t* __restrict__ mask_out,
                                                               int* __restrict__ idx_out, int* __restrict__ cnt_out,
                                                               int* __restrict__ total_out) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.x, n = G * inner;
  const float* lk = logits + (size_t)b * 2 * n;
  const float* ld = lk + n;
  const float* gk = noise ? noise + (size_t)b * 2 * n : nullptr;
  const float* gd = gk ? gk + n : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  int total = 0;
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + threadIdx.x;
    int keep = 0;
    if (i < n) {
      const float a = (lk[i] + (gk ? gk[i] : 0.f)) / tau, d = (ld[i] + (gd ? gd[i] : 0.f)) / tau;
      keep = a >= d ? 1 : 0;
      mask_out[(size_t)b * n + i] = (uint8_t)keep;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (idx_out && inner == 1 && keep) idx_out[(size_t)b * G + before + __popc(bal & ((1u << lane) - 1u))] = i;
    int chunk = 0;
    for (int w = 0; w < 8; ++w) chunk += s_warp[w];
    total += chunk;
    __syncthreads();
    if (threadIdx.x == 0) s_base += chunk;
    __syncthreads();
  }
  if (idx_out && inner == 1) {                       // inactive ids after the active ones, ascending
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 256) {
      const int i = i0 + threadIdx.x;
      const int off = (i < n && mask_out[(size_t)b * n + i] == 0) ? 1 : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, off);
      if (lane == 0) s_warp[warp] = __popc(bal);
      __syncthreads();
      int before = s_base;
      for (int w = 0; w < warp; ++w) before += s_warp[w];
      if (off) idx_out[(size_t)b * G + total + before + __popc(bal & ((1u << lane) - 1u))] = i;
      int chunk = 0;
      for (int w = 0; w < 8; ++w) chunk += s_warp[w];
      __syncthreads();
      if (threadIdx.x == 0) s_base += chunk;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    if (cnt_out) cnt_out[b] = total;
    if (total_out) atomicAdd(total_out, total);
  }
}

extern "C" int laud_gate_from_logits(const float* logits, const float* noise, int B, int G, int inner, float tau,
                                     uint8_t* mask_out, int32_t* idx_out, int32_t* cnt_out, int32_t* total_out, void* stream) {
  LAUD_REQUIRE(logits && mask_out && B > 0 && G > 0 && inner > 0, "laud_gate_from_logits: bad arguments");
  LAUD_REQUIRE(tau > 0.f, "laud_gate_from_logits: temperature must be positive (got %f)", (double)tau);
  LAUD_REQUIRE(!idx_out || inner == 1, "laud_gate_from_logits: index lists exist for channel gates only (inner == 1)");
  gate_from_logits_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(logits, noise, G, inner, tau, mask_out, idx_out, cnt_out, total_out);
  return check_launch("gate_from_logits_kernel");
}
