// Shared helpers for the LAUD sm_100a kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "laud_b200.h"

namespace laud {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
extern std::atomic<unsigned long long> g_conv_paths[3];
extern std::atomic<unsigned long long> g_conv_tma_launches;   // launches of the TMA-staged tcgen05 kernel (subset of g_conv_paths[0])

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LAUD_E_CUDA;
  }
  return LAUD_OK;
}

#define LAUD_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::laud::set_error(__VA_ARGS__);      \
      return LAUD_E_BADARG;                \
    }                                      \
  } while (0)

#define LAUD_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) {                                                \
      ::laud::set_error("%s: %s", #call, cudaGetErrorString(e__));           \
      return LAUD_E_CUDA;                                                    \
    }                                                                        \
  } while (0)

// Per-device one-time state (function attributes, SM count, device tables): a process may drive several GPUs.
constexpr int MAX_DEVICES = 64;
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
  return d;
}
// One-time device-table initialisation kernels are launched on the caller's stream and waited for, so that launches
// on OTHER streams (parallel graph chains) are ordered after them.  Not possible inside a stream capture.
inline int finish_first_call_init(cudaStream_t s, const char* what) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
    set_error("%s: the first call on a device must not be inside a CUDA-graph capture (run one eager forward first)", what);
    return LAUD_E_UNSUPPORTED;
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LAUD_E_CUDA;
  }
  return LAUD_OK;
}

__host__ __device__ __forceinline__ int round_up(int v, int a) { return a > 0 ? (v + a - 1) / a * a : v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Border class of an output coordinate for a 3-tap window: bit0 = first tap
// falls before the input, bit1 = last tap falls past the input.
__device__ __forceinline__ int border_class(int o, int stride, int pad, int in_size) {
  int c = 0;
  if (o * stride - pad < 0) c |= 1;
  if (o * stride + 2 - pad >= in_size) c |= 2;
  return c;
}

// Host-visible view of the conv descriptor with everything the kernels need.
struct ConvArgs {
  const __half* x; int ldx;
  const __half* w;
  __half* y; int ldy;
  int B, H_in, W_in, C_in, H_out, W_out, C_out;
  int ksize, stride, pad;
  const float* scale; const float* shift;
  int relu_mode;
  const __half* residual; int ldr;
  const int* k_idx; const int* k_cnt; int k_ld; int k_gran;
  const int* n_idx; const int* n_cnt; int n_ld; int n_gran;
  const float* pre_bias; int pre_bias_classes; int pre_bias_ld;
  const uint8_t* out_mask; int mask_groups;
  const int* sample_idx; const int* sample_cnt;
  const int* row_idx; const int* row_cnt;
  int n_pad_align;
  float* gap_partial; int gap_tiles;
  const __half* wt;
  const __half* bias_t; int bias_ld;
  const uint8_t* n_mask; int n_mask_gran;
  int n_expand;
  int row_lo, row_hi;       // internal (row lists): this launch handles the list only if row_lo <= *row_cnt < row_hi - the
                            // density-dependent dispatch between the tensor-core and the CUDA-core kernel is decided ON
                            // THE DEVICE (the count never visits the host); 0 / INT_MAX = always
  int gap_hw;               // internal: pixels per sample of the layer (set by the launcher when gap_partial is used)
};

// Per-launch CUDA-event timing of the convolution kernels (bench.py's roofline leg): when enabled through
// laud_conv_profile(), the launch sites bracket the kernel - and only the kernel - with events on its stream.
struct ConvProfScope {
  cudaStream_t s;
  cudaEvent_t e0;
  bool on;
  explicit ConvProfScope(cudaStream_t stream);
  ~ConvProfScope();
};

int conv_forward_naive(const ConvArgs& a, cudaStream_t s);
int conv_forward_hmma(const ConvArgs& a, cudaStream_t s);
int conv_forward_umma(const ConvArgs& a, cudaStream_t s);
bool conv_umma_supported(const ConvArgs& a);
int conv_forward_tma(const ConvArgs& a, cudaStream_t s);
int conv_forward_rows_simt(const ConvArgs& a, cudaStream_t s);   // low-density row lists: vectorised CUDA-core kernel
bool conv_rows_simt_supported(const ConvArgs& a);
bool conv_tma_supported(const ConvArgs& a);

// Shared epilogue: value for output (b, oy, ox), compact channel j / real channel o.
__device__ __forceinline__ float conv_epilogue(const ConvArgs& a, float acc, int b, int oy, int ox,
                                               int j, int o) {
  float v = acc;
  if (a.pre_bias) {
    int cls = 0;
    if (a.pre_bias_classes > 1)
      cls = border_class(oy, a.stride, a.pad, a.H_in) * 4 + border_class(ox, a.stride, a.pad, a.W_in);
    v += a.pre_bias[((size_t)b * a.pre_bias_classes + cls) * a.pre_bias_ld + j];
  }
  if (a.n_mask && a.n_mask[(size_t)b * (a.C_out / a.n_mask_gran) + o / a.n_mask_gran] == 0) v = 0.0f;   // mask before BN
  if (a.scale) v = v * a.scale[o] + a.shift[o];
  uint8_t gate = 1;
  const size_t pix = (size_t)oy * a.W_out + ox;
  if (a.out_mask) {
    int grp = o / (a.C_out / a.mask_groups);
    gate = a.out_mask[((size_t)b * a.mask_groups + grp) * a.H_out * a.W_out + pix];
    if (a.relu_mode != LAUD_RELU_WHERE_GATE0) v = gate ? v : 0.0f;
  }
  if (a.residual)
    v += __half2float(a.residual[((size_t)b * a.H_out * a.W_out + pix) * a.ldr + o]);
  if (a.relu_mode == LAUD_RELU_ALL || (a.relu_mode == LAUD_RELU_WHERE_GATE0 && !gate))
    v = fmaxf(v, 0.0f);
  return v;
}

}  // namespace laud
