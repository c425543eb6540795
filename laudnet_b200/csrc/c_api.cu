// C-ABI glue: error strings, launch accounting, descriptor validation and
// dispatch of the mask-conditioned convolution.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include <vector>
#include "laud_common.cuh"

namespace laud {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_conv_paths[3];
std::atomic<unsigned long long> g_conv_tma_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;

// Inside a stream capture the records become EVENT-RECORD NODES of the graph (cudaEventRecordExternal): every replay
// re-records the same events, so the kernels are timed inside the very graph execution that is being measured.
static void prof_record(cudaEvent_t e, cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone)
    cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);
  else
    cudaEventRecord(e, s);
}
ConvProfScope::ConvProfScope(cudaStream_t stream) : s(stream), e0(nullptr), on(g_prof_on) {
  if (on) {
    cudaEventCreate(&e0);
    prof_record(e0, s);
  }
}
ConvProfScope::~ConvProfScope() {
  if (on) {
    cudaEvent_t e1;
    cudaEventCreate(&e1);
    prof_record(e1, s);
    g_prof_events.emplace_back(e0, e1);
  }
}

}  // namespace laud

using namespace laud;

// 1: start collecting (drops earlier records); 2: stop collecting but KEEP the records (a captured graph keeps
// re-recording them at every replay); 0: stop and drop.
extern "C" void laud_conv_profile(int enable) {
  if (enable == 2) {
    g_prof_on = false;
    return;
  }
  for (auto& pr : g_prof_events) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  g_prof_events.clear();
  g_prof_on = enable != 0;
}
// After a synchronize: device time in ms of every recorded conv launch, in launch order (up to `cap`); returns the
// number of records.
extern "C" int laud_conv_profile_read_all(float* ms_out, int cap) {
  int i = 0;
  for (auto& pr : g_prof_events) {
    float ms = 0.f;
    if (i < cap && ms_out) ms_out[i] = cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess ? ms : -1.f;
    ++i;
  }
  return i;
}
// After a synchronize: number of recorded conv launches and their summed device time in ms.
extern "C" int laud_conv_profile_read(float* total_ms) {
  float tot = 0.f;
  for (auto& pr : g_prof_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) tot += ms;
  }
  if (total_ms) *total_ms = tot;
  return (int)g_prof_events.size();
}

extern "C" int laud_abi_version(void) { return LAUD_ABI_VERSION; }
extern "C" const char* laud_last_error(void) { return g_err; }
extern "C" unsigned long long laud_launch_count(void) { return g_launches.load(); }
extern "C" unsigned long long laud_conv_tma_launch_count(void) { return g_conv_tma_launches.load(); }
extern "C" void laud_conv_path_counts(unsigned long long out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = g_conv_paths[i].load();
}

extern "C" int laud_conv_forward(const laud_conv_desc* d, int impl, void* stream) {
  LAUD_REQUIRE(d != nullptr, "laud_conv_forward: null descriptor");
  LAUD_REQUIRE(d->x && d->w && d->y, "laud_conv_forward: null x/w/y");
  LAUD_REQUIRE(d->B > 0 && d->H_in > 0 && d->W_in > 0 && d->H_out > 0 && d->W_out > 0 && d->C_in > 0 && d->C_out > 0,
               "laud_conv_forward: non-positive dimension");
  LAUD_REQUIRE(d->ksize == 1 || d->ksize == 3, "laud_conv_forward: ksize must be 1 or 3 (got %d)", d->ksize);
  LAUD_REQUIRE(d->stride >= 1 && d->pad >= 0 && d->pad <= 1, "laud_conv_forward: bad stride/pad");
  LAUD_REQUIRE((d->H_in + 2 * d->pad - d->ksize) / d->stride + 1 == d->H_out &&
                   (d->W_in + 2 * d->pad - d->ksize) / d->stride + 1 == d->W_out,
               "laud_conv_forward: output size %dx%d inconsistent with input %dx%d k=%d s=%d p=%d", d->H_out,
               d->W_out, d->H_in, d->W_in, d->ksize, d->stride, d->pad);
  LAUD_REQUIRE(d->C_in % 8 == 0 && d->ldx % 8 == 0 && d->ldy % 2 == 0,
               "laud_conv_forward: need C_in %% 8 == 0, ldx %% 8 == 0, ldy even (C_in=%d ldx=%d ldy=%d)", d->C_in,
               d->ldx, d->ldy);
  LAUD_REQUIRE(d->n_pad_align == 0 || d->n_pad_align == 8 || d->n_pad_align == 16, "laud_conv_forward: n_pad_align");
  LAUD_REQUIRE(d->ldy >= round_up(d->C_out, d->n_pad_align) || d->n_idx, "laud_conv_forward: ldy too small");
  LAUD_REQUIRE((d->scale == nullptr) == (d->shift == nullptr), "laud_conv_forward: scale and shift go together");
  if (d->k_idx) {
    LAUD_REQUIRE(d->k_cnt && d->k_gran >= 1 && d->k_ld * d->k_gran == d->C_in,
                 "laud_conv_forward: k gather needs k_cnt and k_ld*k_gran == C_in");
    LAUD_REQUIRE(d->ldx >= round_up(d->C_in, 8), "laud_conv_forward: ldx too small for compact input");
  } else {
    LAUD_REQUIRE(d->ldx >= d->C_in, "laud_conv_forward: ldx < C_in");
  }
  if (d->n_idx)
    LAUD_REQUIRE(d->n_cnt && d->n_gran >= 1 && d->n_ld * d->n_gran == d->C_out && d->ldy >= round_up(d->C_out, d->n_pad_align),
                 "laud_conv_forward: n gather needs n_cnt, n_ld*n_gran == C_out, ldy >= padded C_out");
  if (d->row_idx) {
    LAUD_REQUIRE(d->row_cnt, "laud_conv_forward: row_idx needs row_cnt");
    LAUD_REQUIRE(!d->k_idx && !d->n_idx && !d->sample_idx && !d->pre_bias,
                 "laud_conv_forward: row gather cannot be combined with per-sample channel/sample lists");
  }
  if (d->sample_idx) LAUD_REQUIRE(d->sample_cnt, "laud_conv_forward: sample_idx needs sample_cnt");
  if (d->pre_bias)
    LAUD_REQUIRE((d->pre_bias_classes == 1 || d->pre_bias_classes == LAUD_PREBIAS_CLASSES) && d->pre_bias_ld > 0,
                 "laud_conv_forward: pre_bias_classes must be 1 or 16");
  if (d->out_mask)
    LAUD_REQUIRE(d->mask_groups >= 1 && d->C_out % d->mask_groups == 0, "laud_conv_forward: bad mask_groups");
  LAUD_REQUIRE(d->relu_mode >= 0 && d->relu_mode <= 2, "laud_conv_forward: bad relu_mode");
  LAUD_REQUIRE(d->relu_mode != LAUD_RELU_WHERE_GATE0 || d->out_mask, "laud_conv_forward: RELU_WHERE_GATE0 needs out_mask");
  if (d->residual) LAUD_REQUIRE(d->ldr >= d->C_out && !d->n_idx, "laud_conv_forward: residual needs dense output channels");
  if (d->gap_partial)
    LAUD_REQUIRE(d->gap_tiles >= 1 && (impl == LAUD_CONV_AUTO || impl == LAUD_CONV_UMMA),
                 "laud_conv_forward: gap_partial needs gap_tiles >= 1 and the tcgen05 path");
  if (d->bias_t)
    LAUD_REQUIRE(d->w_t && d->k_idx && !d->pre_bias && d->bias_ld % 8 == 0 && d->ksize * d->ksize <= 9 &&
                     (reinterpret_cast<uintptr_t>(d->bias_t) & 15) == 0,
                 "laud_conv_forward: bias_t needs the w_t path (k_idx + w_t), no pre_bias, bias_ld %% 8 == 0");
  if (d->n_mask)
    LAUD_REQUIRE(!d->n_idx && !d->k_idx && !d->pre_bias && d->n_mask_gran >= 1 && d->C_out % d->n_mask_gran == 0,
                 "laud_conv_forward: n_mask (masked-dense gate) excludes n_idx / k_idx / pre_bias and needs C_out %% n_mask_gran == 0");
  if (d->n_expand) {
    LAUD_REQUIRE(d->n_idx && d->n_cnt && !d->k_idx && !d->residual && !d->out_mask && !d->sample_idx && !d->row_idx &&
                     !d->pre_bias && !d->bias_t && !d->gap_partial && d->scale && (d->n_gran & 1) == 0 &&
                     d->relu_mode != LAUD_RELU_WHERE_GATE0 && d->ldy >= d->C_out,
                 "laud_conv_forward: n_expand needs n_idx/n_cnt with an even granularity, scale/shift, ldy >= C_out and excludes "
                 "k_idx / residual / out_mask / lists / pre_bias / bias_t / gap_partial");
    LAUD_REQUIRE(impl == LAUD_CONV_AUTO || impl == LAUD_CONV_UMMA, "laud_conv_forward: n_expand is a tcgen05-path argument");
  }
  if (d->w_t && d->k_idx)
    LAUD_REQUIRE(impl == LAUD_CONV_AUTO || impl == LAUD_CONV_UMMA,
                 "laud_conv_forward: w_t (K-row-gather path, real-channel pre_bias) is a tcgen05-path argument");

  ConvArgs a;
  a.x = (const __half*)d->x; a.ldx = d->ldx;
  a.w = (const __half*)d->w;
  a.y = (__half*)d->y; a.ldy = d->ldy;
  a.B = d->B; a.H_in = d->H_in; a.W_in = d->W_in; a.C_in = d->C_in;
  a.H_out = d->H_out; a.W_out = d->W_out; a.C_out = d->C_out;
  a.ksize = d->ksize; a.stride = d->stride; a.pad = d->pad;
  a.scale = d->scale; a.shift = d->shift; a.relu_mode = d->relu_mode;
  a.residual = (const __half*)d->residual; a.ldr = d->ldr;
  a.k_idx = d->k_idx; a.k_cnt = d->k_cnt; a.k_ld = d->k_ld; a.k_gran = d->k_gran;
  a.n_idx = d->n_idx; a.n_cnt = d->n_cnt; a.n_ld = d->n_ld; a.n_gran = d->n_gran;
  a.pre_bias = d->pre_bias; a.pre_bias_classes = d->pre_bias_classes; a.pre_bias_ld = d->pre_bias_ld;
  a.out_mask = d->out_mask; a.mask_groups = d->mask_groups;
  a.sample_idx = d->sample_idx; a.sample_cnt = d->sample_cnt;
  a.row_idx = d->row_idx; a.row_cnt = d->row_cnt;
  a.n_pad_align = d->n_pad_align;
  a.gap_partial = d->gap_partial; a.gap_tiles = d->gap_tiles;
  a.wt = (const __half*)d->w_t;
  a.bias_t = (const __half*)d->bias_t; a.bias_ld = d->bias_ld;
  a.n_mask = d->n_mask; a.n_mask_gran = d->n_mask_gran;
  a.n_expand = d->n_expand;
  a.row_lo = 0;
  a.row_hi = 0x7fffffff;
  a.gap_hw = 0;

  cudaStream_t s = (cudaStream_t)stream;
  switch (impl) {
    case LAUD_CONV_AUTO:
    case LAUD_CONV_UMMA:
      if (a.wt && a.k_idx && !conv_umma_supported(a)) {
        set_error("laud_conv_forward: layout not supported by the tcgen05 kernel but w_t was given "
                  "(channel granularity must be 2, 4 or a multiple of 8; pitches multiples of 8; 16-byte aligned)");
        return LAUD_E_UNSUPPORTED;
      }
      if (a.row_idx && impl == LAUD_CONV_AUTO && conv_rows_simt_supported(a) && conv_umma_supported(a)) {
        // Row lists (spatial skipping): density-dependent dispatch ON THE DEVICE.  Both kernels are enqueued; each reads
        // *row_cnt and returns at once unless the count is in its regime - below LAUD_ROWS_SIMT_MAX active pixels (less
        // than one 128-row MMA tile) the vectorised CUDA-core kernel, above it the tcgen05 kernel.
        static const int simt_max = getenv("LAUD_ROWS_SIMT_MAX") ? atoi(getenv("LAUD_ROWS_SIMT_MAX")) : 96;
        if (simt_max > 0) {
          ConvArgs lo = a, hi = a;
          lo.row_hi = simt_max;
          hi.row_lo = simt_max;
          if (int e = conv_forward_rows_simt(lo, s)) return e;
          return conv_forward_umma(hi, s);
        }
      }
      {
        static const bool force_v3 = getenv("LAUD_CONV_V3") != nullptr;     // A/B switch for profiling
        if (!force_v3 && conv_tma_supported(a)) return conv_forward_tma(a, s);
      }
      if (a.gap_partial) {
        set_error("laud_conv_forward: fused GAP (gap_partial) is only available on the TMA-staged kernel");
        return LAUD_E_UNSUPPORTED;
      }
      if (a.n_expand) {
        set_error("laud_conv_forward: n_expand (gather4 channel skipping) needs a 3x3 stride-1 layer with W_out + 2 <= 128 on the "
                  "TMA-staged kernel");
        return LAUD_E_UNSUPPORTED;
      }
      return conv_forward_umma(a, s);
    case LAUD_CONV_HMMA:
      return conv_forward_hmma(a, s);
    case LAUD_CONV_NAIVE:
      return conv_forward_naive(a, s);
    default:
      set_error("laud_conv_forward: unknown impl %d", impl);
      return LAUD_E_BADARG;
  }
}
