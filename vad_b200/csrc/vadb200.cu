// vadb200.cu -- C ABI (include/vadb200.h) over the sm_100a kernels of vad_kernels.cuh.
// Host logic only: handle / plan / bank lifetime, segment tables, launches, the chunked
// H2D -> kernel -> D2H pipeline of vadb200_vad_host.  No CPU compute path exists: every
// numeric entry point launches a CUDA kernel or fails with VADB200_E_CUDA.
#include "../../include/vadb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "vad_host_tables.h"
#include "vad_kernels.cuh"
#include "ffn_train.cuh"

using namespace vadb;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<unsigned long long> g_next_id{1};
std::mutex g_tab_mu;
bool g_tab_loaded[64] = {false};  // per device: mel / DCT tables uploaded (config-invariant content)

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return VADB200_E_CUDA;
}
#define CU(expr)                                     \
  do {                                               \
    cudaError_t e__ = (expr);                        \
    if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
  } while (0)

long long* g_dbg_ts = nullptr;  // timing experiments only (vadb200_debug_timestamps)
constexpr int kNumSmFallback = 148;
constexpr int kPlanCounters = 16;  // launches of one plan that may be in flight at once (any streams)
constexpr int kHostBufs = 3;

}  // namespace

struct vadb200_handle {
  int device = 0;
  unsigned long long id = 0;
  unsigned version = 1;
  MfccConfig cfg;
  std::vector<double> fb;
  MelDctTables tab;   // identical for every handle (only the reference configuration is compiled in)
  FfnParams par;      // per-handle FFN weights: passed to kernels as a launch parameter
  FfnBias bias;       // the same biases for the tensor-core variant
  int plan_seg_frames = 0;  // 0 = automatic segment length (vadb200_set_plan_segment_frames)
  bool have_ffn = false;
  cf2* d_tw = nullptr;  // tw1[256] then tw2[128]
  int num_sms = kNumSmFallback;
  bool attr_set = false;
  long long host_chunk_samples = 32ll << 20;
  // vad_host pipeline resources (lazy)
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kHostBufs] = {}, ev_done[kHostBufs] = {}, ev_out[kHostBufs] = {};
  int16_t* d_stage[kHostBufs] = {};
  uint8_t* d_lab[kHostBufs] = {};
  float* d_lgt[kHostBufs] = {};
  float* d_rows[kHostBufs] = {};
  long long stage_cap = 0, lab_cap = 0, lgt_cap = 0, rows_cap = 0;
  float* d_sink = nullptr;
  unsigned char* d_tc_blob = nullptr;   // canonical tf32 hi/lo weight blob (ffn_tc.cuh)
  unsigned char* d_tc16_blob = nullptr; // canonical fp16 hi/lo weight blob, statically scaled
  bool tc16_ok = false;                 // the fp16 scales are in range for this handle's weights
  int ffn_impl = 2;                     // 0: FP32 CUDA cores, 1: tcgen05 tf32 hi/lo, 2: tcgen05 fp16 hi/lo (default)
};

struct vadb200_plan {
  vadb200_handle* h = nullptr;
  int mode = 0;
  long long n_utt = 0;
  std::vector<long long> offsets, lengths, row_off, utt_seg;  // utt_seg: n_utt + 1
  std::vector<Segment> segs;
  Segment* d_segs = nullptr;
  int* d_counter = nullptr;   // kPlanCounters self-resetting (next segment, CTAs done) pairs
  std::atomic<unsigned> launch_seq{0};
  long long total_rows = 0;
  long long seg_frames = 0;
};

struct vadb200_trainer {
  vadb200_handle* h = nullptr;
  long long max_batch = 0;
  float lr = 1.0f, rho = 0.95f, eps = 1e-8f;
  float* d_state = nullptr;    // params | acc_g | acc_u   (3 x kNParams)
  float* d_partial = nullptr;  // [min(ceil(max_batch / 128), SMs)][kNParams + 1]
  float* d_loss = nullptr;
  long long steps = 0;
};

struct vadb200_bank {
  vadb200_handle* h = nullptr;
  int n = 0;
  int16_t* d_hist = nullptr;
  float* d_ring = nullptr;
  int* d_fed = nullptr;
};

namespace {

// Mel weights and the folded DCT matrix are the same for every handle (the kernels are compiled for the
// reference configuration only), so the __constant__ copy is written once per device and never changes:
// no handle can observe another handle's state through it.  FFN weights never go through device globals.
int ensure_tables(vadb200_handle* h) {
  if (h->device >= 64) return fail(VADB200_E_INVALID, "device index >= 64");
  std::lock_guard<std::mutex> lk(g_tab_mu);
  if (g_tab_loaded[h->device]) return 0;
  CU(cudaMemcpyToSymbol(c_tab, &h->tab, sizeof(MelDctTables), 0, cudaMemcpyHostToDevice));
  CU(cudaDeviceSynchronize());
  g_tab_loaded[h->device] = true;
  return 0;
}

int ensure_attrs(vadb200_handle* h) {
  if (h->attr_set) return 0;
  CU(cudaFuncSetAttribute(fused_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
  CU(cudaFuncSetAttribute(fused_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
  CU(cudaFuncSetAttribute(fused_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
  CU(cudaFuncSetAttribute(fused_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
  CU(cudaFuncSetAttribute(fused_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemBytes));
  CU(cudaFuncSetAttribute(ffn_tc_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnTcSmemBytes));
  CU(cudaFuncSetAttribute(ffn_tc_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnTcSmemBytes));
  CU(cudaFuncSetAttribute(frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFramesSmemBytes));
  CU(cudaFuncSetAttribute(stream_feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBytes));
  h->attr_set = true;
  return 0;
}

int launch_fused(vadb200_plan* p, FusedParams fp, int n_segs, cudaStream_t st) {
  vadb200_handle* h = p->h;
  if (n_segs <= 0) return 0;
  int rc = ensure_attrs(h);
  if (rc) return rc;
  rc = ensure_tables(h);
  if (rc) return rc;
  // each launch takes its own self-resetting counter pair, so one plan can be in flight on several streams
  fp.counter = p->d_counter + 2 * (p->launch_seq.fetch_add(1) % kPlanCounters);
#if defined(VADB_DEBUG_HOOKS)
  const char* gps = std::getenv("VADB200_CTAS_PER_SM");  // occupancy experiments only
  const int per_sm = gps ? std::max(1, std::atoi(gps)) : 2;
#else
  const int per_sm = 2;
#endif
  const int grid = std::min(n_segs, per_sm * h->num_sms);
  switch (p->mode) {
    case VADB200_MODE_MFCC: fused_kernel<0, 0><<<grid, kThreads, kFusedSmemBytes, st>>>(fp, FfnNone{0}); break;
    case VADB200_MODE_DATASET: fused_kernel<1, 0><<<grid, kThreads, kFusedSmemBytes, st>>>(fp, FfnNone{0}); break;
    default:
      if (h->ffn_impl == 2 && h->tc16_ok) {
        fp.tc_blob = h->d_tc16_blob;
        fused_kernel<2, 2><<<grid, kThreads, kFusedSmemBytes, st>>>(fp, h->bias);
      } else if (h->ffn_impl >= 1) {
        fused_kernel<2, 1><<<grid, kThreads, kFusedSmemBytes, st>>>(fp, h->bias);
      } else {
        fused_kernel<2, 0><<<grid, kThreads, kFusedSmemBytes, st>>>(fp, h->par);
      }
      break;
  }
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

FusedParams base_params(vadb200_plan* p) {
  FusedParams fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.segs = p->d_segs;
  fp.counter = p->d_counter;  // launch_fused picks the launch's own pair
  fp.tw1 = p->h->d_tw;
  fp.tw2 = p->h->d_tw + 256;
  fp.tc_blob = p->h->d_tc_blob;
#if defined(VADB_DEBUG_HOOKS)
  const char* dbg = std::getenv("VADB200_DEBUG_SKIP");  // timing experiments only; results are garbage
  fp.debug_skip = dbg ? std::atoi(dbg) : 0;
  fp.dbg_ts = g_dbg_ts;
#else
  fp.debug_skip = 0;
  fp.dbg_ts = nullptr;
#endif
  return fp;
}

}  // namespace

extern "C" {

const char* vadb200_last_error(void) { return g_err.c_str(); }
int vadb200_version(void) { return VADB200_VERSION; }
int64_t vadb200_launch_count(void) { return g_launches.load(); }
// Timing experiments only (not in the public header): device buffer of 64 x 16 clock64 stamps written by
// CTA 0 during its first block phases; pass NULL to switch off.
int vadb200_debug_timestamps(long long* d_buf) { g_dbg_ts = d_buf; return 0; }

void vadb200_default_config(vadb200_config* c) {
  if (!c) return;
  c->sample_rate = 16000; c->frame_size = 400; c->frame_step = 160; c->fft_n = 512;
  c->n_filters = 26; c->n_mfcc = 13; c->low_hz = 300.0; c->high_hz = 8000.0; c->lifter_l = 22;
  c->reserved = 0;
}

int64_t vadb200_frames_for_length(int64_t n) { return frames_for_length(n); }
int64_t vadb200_outputs_for_length(int64_t n) { return outputs_for_length(n); }

int vadb200_create(const vadb200_config* c, int device, vadb200_handle** out) {
  if (!out) return fail(VADB200_E_INVALID, "out is null");
  *out = nullptr;
  vadb200_config dc;
  if (!c) {
    vadb200_default_config(&dc);
    c = &dc;
  }
  MfccConfig cfg;
  cfg.sample_rate = c->sample_rate; cfg.frame_size = c->frame_size; cfg.frame_step = c->frame_step;
  cfg.fft_n = c->fft_n; cfg.n_filters = c->n_filters; cfg.n_mfcc = c->n_mfcc;
  cfg.low_hz = c->low_hz; cfg.high_hz = c->high_hz; cfg.lifter_l = c->lifter_l;
  if (!is_reference_config(cfg))
    return fail(VADB200_E_UNSUPPORTED,
                "only the reference configuration (16 kHz, 400/160, n_fft 512, 26 filters 300-8000 Hz, 13 "
                "cepstra, lifter 22; config.py:20-27) is compiled in");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(VADB200_E_INVALID, "no such CUDA device");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(VADB200_E_UNSUPPORTED, "kernels are built for sm_100a (B200) only");
  vadb200_handle* h = new (std::nothrow) vadb200_handle();
  if (!h) return fail(VADB200_E_NOMEM, "host allocation failed");
  h->device = device;
  h->id = g_next_id.fetch_add(1);
  h->cfg = cfg;
  h->num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : kNumSmFallback;
  h->fb = mel_filterbank(cfg);
  std::memset(&h->par, 0, sizeof(FfnParams));
  std::memset(&h->bias, 0, sizeof(FfnBias));
  std::memset(&h->tab, 0, sizeof(MelDctTables));
  std::string why;
  if (!pack_mel_weights(h->fb.data(), h->tab.melw, &why)) {
    delete h;
    return fail(VADB200_E_UNSUPPORTED, why);
  }
  folded_dct(cfg, h->tab.dct);
  pack_mel_pairs(h->fb.data(), h->tab.melw2);
  pack_dct_pairs(h->tab.dct, h->tab.dctp);
  cf2 tw[384];
  fft_twiddles(tw, tw + 256);
  cudaError_t e = cudaMalloc(&h->d_tw, sizeof(tw));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_tw, tw, sizeof(tw), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_sink, 256);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_tc_blob, kTcBlobBytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_tc16_blob, kTc16BlobBytes);
  if (e != cudaSuccess) {
    cudaFree(h->d_tw);
    cudaFree(h->d_sink);
    cudaFree(h->d_tc_blob);
    delete h;
    return cuda_fail(e, "vadb200_create");
  }
  // function attributes and the (config-invariant) constant tables now, so that no launch ever
  // synchronises: every device-pointer entry point can be captured into a CUDA graph
  int rc = ensure_attrs(h);
  if (!rc) rc = ensure_tables(h);
  if (rc) {
    const std::string keep = g_err;
    vadb200_destroy(h);
    g_err = keep;
    return rc;
  }
  *out = h;
  return 0;
}

int vadb200_destroy(vadb200_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (int i = 0; i < kHostBufs; ++i) {
    cudaFree(h->d_stage[i]); cudaFree(h->d_lab[i]); cudaFree(h->d_lgt[i]); cudaFree(h->d_rows[i]);
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
  }
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_run) cudaStreamDestroy(h->s_run);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  cudaFree(h->d_tw);
  cudaFree(h->d_sink);
  cudaFree(h->d_tc_blob);
  cudaFree(h->d_tc16_blob);
  delete h;
  return 0;
}

int vadb200_get_filterbank(vadb200_handle* h, double* out) {
  if (!h || !out) return fail(VADB200_E_INVALID, "null argument");
  std::memcpy(out, h->fb.data(), h->fb.size() * sizeof(double));
  return 0;
}

int vadb200_set_ffn_weights(vadb200_handle* h, const float* W1, const float* b1, const float* W2, const float* b2,
                            const float* W3, const float* b3, const float* W4, const float* b4) {
  if (!h || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !W4 || !b4) return fail(VADB200_E_INVALID, "null argument");
  std::memcpy(h->par.W1, W1, sizeof(h->par.W1)); std::memcpy(h->par.b1, b1, sizeof(h->par.b1));
  std::memcpy(h->par.W2, W2, sizeof(h->par.W2)); std::memcpy(h->par.b2, b2, sizeof(h->par.b2));
  std::memcpy(h->par.W3, W3, sizeof(h->par.W3)); std::memcpy(h->par.b3, b3, sizeof(h->par.b3));
  std::memcpy(h->par.W4, W4, sizeof(h->par.W4)); std::memcpy(h->par.b4, b4, sizeof(h->par.b4));
  std::memcpy(h->bias.b1, b1, sizeof(h->bias.b1)); std::memcpy(h->bias.b2, b2, sizeof(h->bias.b2));
  std::memcpy(h->bias.b3, b3, sizeof(h->bias.b3)); std::memcpy(h->bias.b4, b4, sizeof(h->bias.b4));
  h->have_ffn = true;
  h->version = (h->version + 1) & 0xFFFFF;
  // tensor-core operand blob; drained + blocking so kernels on any stream see a consistent copy
  std::vector<unsigned char> blob(kTcBlobBytes);
  tc_pack_weights(h->par, blob.data());
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(h->d_tc_blob, blob.data(), kTcBlobBytes, cudaMemcpyHostToDevice));
  std::vector<unsigned char> blob16(kTc16BlobBytes);
  h->tc16_ok = tc16_pack_weights(h->par, blob16.data(), h->bias);
  if (h->tc16_ok) CU(cudaMemcpy(h->d_tc16_blob, blob16.data(), kTc16BlobBytes, cudaMemcpyHostToDevice));
  return 0;
}

int vadb200_set_ffn_impl(vadb200_handle* h, int impl) {
  if (!h || impl < 0 || impl > 2)
    return fail(VADB200_E_INVALID, "impl must be 0 (fp32), 1 (tcgen05 tf32 hi/lo) or 2 (tcgen05 fp16 hi/lo)");
  h->ffn_impl = impl;
  return 0;
}
int vadb200_get_ffn_impl(vadb200_handle* h) { return h ? h->ffn_impl : -1; }

int vadb200_set_host_chunk_samples(vadb200_handle* h, int64_t samples) {
  if (!h || samples < 4096) return fail(VADB200_E_INVALID, "chunk must be >= 4096 samples");
  h->host_chunk_samples = samples & ~7ll;
  return 0;
}

// ---- plans ---------------------------------------------------------------------------------------------
int vadb200_plan_create(vadb200_handle* h, const int64_t* offs, const int64_t* lens, int64_t n_utt, int mode,
                        vadb200_plan** out) {
  if (!h || !out || n_utt < 0 || (n_utt > 0 && (!offs || !lens))) return fail(VADB200_E_INVALID, "null / negative argument");
  if (mode < 0 || mode > 2) return fail(VADB200_E_INVALID, "unknown mode");
  *out = nullptr;
  vadb200_plan* p = new (std::nothrow) vadb200_plan();
  if (!p) return fail(VADB200_E_NOMEM, "host allocation failed");
  p->h = h; p->mode = mode; p->n_utt = n_utt;
  p->offsets.assign(offs, offs + n_utt);
  p->lengths.assign(lens, lens + n_utt);
  p->row_off.resize(n_utt + 1);
  p->utt_seg.resize(n_utt + 1);
  long long rows = 0, prev_end = 0;
  for (long long u = 0; u < n_utt; ++u) {
    if (offs[u] < 0 || (offs[u] & 7) || lens[u] < 0 || offs[u] < prev_end) {
      delete p;
      return fail(VADB200_E_INVALID, "utterance offsets must be ascending, non-overlapping multiples of 8 samples");
    }
    prev_end = offs[u] + lens[u];
    p->row_off[u] = rows;
    rows += (mode == VADB200_MODE_MFCC) ? frames_for_length(lens[u]) : outputs_for_length(lens[u]);
  }
  p->row_off[n_utt] = rows;
  p->total_rows = rows;
  // Segment length: long runs amortise the 4-frame halo; short ones keep all CTAs busy on small
  // batches.  n_frames of a full segment is a multiple of 32 (whole steps).
  const int halo = (mode == VADB200_MODE_MFCC) ? 0 : 4;
  long long seg_frames = h->plan_seg_frames;
  if (seg_frames <= 0) {
    // >= 16 segments per persistent CTA when the batch allows (dynamic scheduling then balances to a few
    // per cent), whole 128-frame tensor-core tiles, at most 2048 frames
    const long long want = rows / (16ll * 2 * h->num_sms) + halo;
    seg_frames = want <= 64 ? 64 : std::min<long long>(2048, ((want + 127) / 128) * 128);
  }
  p->seg_frames = seg_frames;
  const long long seg_rows = seg_frames - halo;
  for (long long u = 0; u < n_utt; ++u) {
    p->utt_seg[u] = static_cast<long long>(p->segs.size());
    const long long r = p->row_off[u + 1] - p->row_off[u];
    for (long long o = 0; o < r; o += seg_rows) {
      Segment s;
      s.pcm_start = offs[u] + kHop * o;
      s.out_start = p->row_off[u] + o;
      s.n_frames = static_cast<int>(std::min(seg_rows, r - o) + halo);
      s.pad = 0;
      p->segs.push_back(s);
    }
  }
  p->utt_seg[n_utt] = static_cast<long long>(p->segs.size());
  if (p->segs.size() > 0x7fffffffull) {
    delete p;
    return fail(VADB200_E_INVALID, "too many segments");
  }
  cudaError_t e = cudaSetDevice(h->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_counter, 2 * kPlanCounters * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(p->d_counter, 0, 2 * kPlanCounters * sizeof(int));
  if (e == cudaSuccess && !p->segs.empty()) {
    e = cudaMalloc(&p->d_segs, p->segs.size() * sizeof(Segment));
    if (e == cudaSuccess)
      e = cudaMemcpy(p->d_segs, p->segs.data(), p->segs.size() * sizeof(Segment), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    cudaFree(p->d_counter); cudaFree(p->d_segs);
    delete p;
    return cuda_fail(e, "vadb200_plan_create");
  }
  *out = p;
  return 0;
}

int vadb200_plan_destroy(vadb200_plan* p) {
  if (!p) return 0;
  cudaSetDevice(p->h->device);
  cudaFree(p->d_segs);
  cudaFree(p->d_counter);
  delete p;
  return 0;
}

int64_t vadb200_plan_total_rows(const vadb200_plan* p) { return p ? p->total_rows : -1; }
int64_t vadb200_plan_segment_frames(const vadb200_plan* p) { return p ? p->seg_frames : -1; }
int64_t vadb200_plan_segment_count(const vadb200_plan* p) { return p ? static_cast<int64_t>(p->segs.size()) : -1; }

int vadb200_set_plan_segment_frames(vadb200_handle* h, int frames) {
  if (!h || frames < 0 || frames > 2048 || (frames != 0 && (frames < 64 || frames % 32)))
    return fail(VADB200_E_INVALID, "segment length must be 0 (automatic) or a multiple of 32 in [64, 2048]");
  h->plan_seg_frames = frames;
  return 0;
}

int vadb200_plan_row_offsets(const vadb200_plan* p, int64_t* out) {
  if (!p || !out) return fail(VADB200_E_INVALID, "null argument");
  for (long long u = 0; u <= p->n_utt; ++u) out[u] = p->row_off[u];
  return 0;
}

static int check_pcm(const vadb200_plan* p, const void* pcm, int64_t pcm_len) {
  if (!p) return fail(VADB200_E_INVALID, "plan is null");
  if (p->total_rows > 0 && !pcm) return fail(VADB200_E_INVALID, "pcm is null");
  if (reinterpret_cast<uintptr_t>(pcm) & 15) return fail(VADB200_E_INVALID, "pcm must be 16-byte aligned");
  if (p->n_utt > 0 && pcm_len < p->offsets[p->n_utt - 1] + p->lengths[p->n_utt - 1])
    return fail(VADB200_E_INVALID, "pcm_len is smaller than the last utterance's end");
  return 0;
}

int vadb200_mfcc_packed(vadb200_plan* p, const int16_t* d_pcm, int64_t pcm_len, float* d_out, void* stream) {
  int rc = check_pcm(p, d_pcm, pcm_len);
  if (rc) return rc;
  if (p->mode == VADB200_MODE_VAD) return fail(VADB200_E_INVALID, "plan was created for MODE_VAD");
  if (p->total_rows > 0 && !d_out) return fail(VADB200_E_INVALID, "d_out is null");
  CU(cudaSetDevice(p->h->device));
  FusedParams fp = base_params(p);
  fp.pcm = d_pcm; fp.pcm_len = pcm_len; fp.rows = d_out;
  fp.seg_begin = 0; fp.seg_end = static_cast<int>(p->segs.size());
  return launch_fused(p, fp, fp.seg_end, static_cast<cudaStream_t>(stream));
}

int vadb200_vad_packed(vadb200_plan* p, const int16_t* d_pcm, int64_t pcm_len, uint8_t* d_labels, float* d_logits,
                       float* d_feats, int feat_mode, void* stream) {
  int rc = check_pcm(p, d_pcm, pcm_len);
  if (rc) return rc;
  if (p->mode != VADB200_MODE_VAD) return fail(VADB200_E_INVALID, "plan was not created for MODE_VAD");
  if (!p->h->have_ffn) return fail(VADB200_E_STATE, "FFN weights not set (vadb200_set_ffn_weights)");
  if (p->total_rows > 0 && !d_labels) return fail(VADB200_E_INVALID, "d_labels is null");
  if (feat_mode != VADB200_FEAT_ANALYSER && feat_mode != VADB200_FEAT_DATASET) return fail(VADB200_E_INVALID, "unknown feat_mode");
  CU(cudaSetDevice(p->h->device));
  FusedParams fp = base_params(p);
  fp.pcm = d_pcm; fp.pcm_len = pcm_len; fp.labels = d_labels; fp.logits = d_logits; fp.feats = d_feats;
  fp.feat_mode = feat_mode;
  fp.seg_begin = 0; fp.seg_end = static_cast<int>(p->segs.size());
  return launch_fused(p, fp, fp.seg_end, static_cast<cudaStream_t>(stream));
}

// ---- host-buffer pipeline ---------------------------------------------------------------------------------
static int ensure_host_pipeline(vadb200_handle* h, long long stage_samples, long long rows, bool want_labels,
                                bool want_logits, long long row_floats) {
  if (!h->s_in) {
    CU(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->s_run, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < kHostBufs; ++i) {
      CU(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
    }
  }
  if (stage_samples > h->stage_cap) {
    for (int i = 0; i < kHostBufs; ++i) {
      cudaFree(h->d_stage[i]); h->d_stage[i] = nullptr;
      CU(cudaMalloc(&h->d_stage[i], stage_samples * sizeof(int16_t)));
    }
    h->stage_cap = stage_samples;
  }
  if (want_labels && rows > h->lab_cap) {
    for (int i = 0; i < kHostBufs; ++i) {
      cudaFree(h->d_lab[i]); h->d_lab[i] = nullptr;
      CU(cudaMalloc(&h->d_lab[i], rows));
    }
    h->lab_cap = rows;
  }
  if (want_logits && rows > h->lgt_cap) {
    for (int i = 0; i < kHostBufs; ++i) {
      cudaFree(h->d_lgt[i]); h->d_lgt[i] = nullptr;
      CU(cudaMalloc(&h->d_lgt[i], rows * 3 * sizeof(float)));
    }
    h->lgt_cap = rows;
  }
  if (row_floats > h->rows_cap) {
    for (int i = 0; i < kHostBufs; ++i) {
      cudaFree(h->d_rows[i]); h->d_rows[i] = nullptr;
      CU(cudaMalloc(&h->d_rows[i], row_floats * sizeof(float)));
    }
    h->rows_cap = row_floats;
  }
  return 0;
}

// Chunked H2D -> fused kernel -> D2H over three staging buffers and three streams.  VAD plans return
// labels (+ logits); MFCC / dataset plans return float rows of `width` columns.
static int run_host_pipeline(vadb200_plan* p, const int16_t* h_pcm, int64_t pcm_len, uint8_t* h_labels, float* h_logits,
                             float* h_rows, int feat_mode) {
  vadb200_handle* h = p->h;
  if (p->total_rows == 0) return 0;
  if (p->n_utt > 0 && pcm_len < p->offsets[p->n_utt - 1] + p->lengths[p->n_utt - 1])
    return fail(VADB200_E_INVALID, "pcm_len is smaller than the last utterance's end");
  CU(cudaSetDevice(h->device));
  const long long width = p->mode == VADB200_MODE_VAD ? 0 : (p->mode == VADB200_MODE_MFCC ? kNCep : kNFeat);
  // chunks = runs of whole utterances whose PCM span fits host_chunk_samples
  struct Chunk { long long u0, u1, s0, s1; };
  std::vector<Chunk> chunks;
  long long max_span = 0, max_rows = 0;
  for (long long u = 0; u < p->n_utt;) {
    Chunk c{u, u, p->offsets[u], 0};
    while (c.u1 < p->n_utt && (c.u1 == c.u0 || p->offsets[c.u1] + p->lengths[c.u1] - c.s0 <= h->host_chunk_samples)) ++c.u1;
    c.s1 = p->offsets[c.u1 - 1] + p->lengths[c.u1 - 1];
    max_span = std::max(max_span, c.s1 - c.s0);
    max_rows = std::max(max_rows, p->row_off[c.u1] - p->row_off[c.u0]);
    chunks.push_back(c);
    u = c.u1;
  }
  max_rows = std::max<long long>(max_rows, 1);
  int rc = ensure_host_pipeline(h, ((max_span + 7) & ~7ll) + 8, max_rows, h_labels != nullptr, h_logits != nullptr,
                                width * max_rows);
  if (rc) return rc;
  for (size_t i = 0; i < chunks.size(); ++i) {
    const Chunk& c = chunks[i];
    const int b = static_cast<int>(i % kHostBufs);
    const long long span = c.s1 - c.s0;
    const long long r0 = p->row_off[c.u0], nrows = p->row_off[c.u1] - r0;
    if (i >= kHostBufs) CU(cudaStreamWaitEvent(h->s_in, h->ev_done[b], 0));  // stage buffer free again
    CU(cudaMemcpyAsync(h->d_stage[b], h_pcm + c.s0, span * sizeof(int16_t), cudaMemcpyHostToDevice, h->s_in));
    CU(cudaEventRecord(h->ev_in[b], h->s_in));
    CU(cudaStreamWaitEvent(h->s_run, h->ev_in[b], 0));
    if (i >= kHostBufs) CU(cudaStreamWaitEvent(h->s_run, h->ev_out[b], 0));  // result buffers drained
    FusedParams fp = base_params(p);
    fp.pcm = h->d_stage[b] - c.s0;  // segment sample indices stay absolute
    fp.pcm_len = c.s1;
    fp.labels = h_labels ? h->d_lab[b] : nullptr;
    fp.logits = h_logits ? h->d_lgt[b] : nullptr;
    fp.rows = width ? h->d_rows[b] : nullptr;
    fp.row_base = r0;
    fp.feat_mode = feat_mode;
    fp.seg_begin = static_cast<int>(p->utt_seg[c.u0]);
    fp.seg_end = static_cast<int>(p->utt_seg[c.u1]);
    rc = launch_fused(p, fp, fp.seg_end - fp.seg_begin, h->s_run);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev_done[b], h->s_run));
    CU(cudaStreamWaitEvent(h->s_out, h->ev_done[b], 0));
    if (nrows > 0) {
      if (h_labels) CU(cudaMemcpyAsync(h_labels + r0, h->d_lab[b], nrows, cudaMemcpyDeviceToHost, h->s_out));
      if (h_logits)
        CU(cudaMemcpyAsync(h_logits + r0 * 3, h->d_lgt[b], nrows * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
      if (width)
        CU(cudaMemcpyAsync(h_rows + r0 * width, h->d_rows[b], nrows * width * sizeof(float), cudaMemcpyDeviceToHost,
                           h->s_out));
    }
    CU(cudaEventRecord(h->ev_out[b], h->s_out));
  }
  CU(cudaStreamSynchronize(h->s_out));
  CU(cudaStreamSynchronize(h->s_run));
  return 0;
}

int vadb200_vad_host(vadb200_plan* p, const int16_t* h_pcm, int64_t pcm_len, uint8_t* h_labels, float* h_logits,
                     int feat_mode) {
  if (!p) return fail(VADB200_E_INVALID, "plan is null");
  if (p->mode != VADB200_MODE_VAD) return fail(VADB200_E_INVALID, "plan was not created for MODE_VAD");
  if (!p->h->have_ffn) return fail(VADB200_E_STATE, "FFN weights not set (vadb200_set_ffn_weights)");
  if (feat_mode != VADB200_FEAT_ANALYSER && feat_mode != VADB200_FEAT_DATASET) return fail(VADB200_E_INVALID, "unknown feat_mode");
  if (p->total_rows == 0) return 0;
  if (!h_pcm || !h_labels) return fail(VADB200_E_INVALID, "null argument");
  return run_host_pipeline(p, h_pcm, pcm_len, h_labels, h_logits, nullptr, feat_mode);
}

int vadb200_mfcc_host(vadb200_plan* p, const int16_t* h_pcm, int64_t pcm_len, float* h_out) {
  if (!p) return fail(VADB200_E_INVALID, "plan is null");
  if (p->mode == VADB200_MODE_VAD) return fail(VADB200_E_INVALID, "plan was created for MODE_VAD");
  if (p->total_rows == 0) return 0;
  if (!h_pcm || !h_out) return fail(VADB200_E_INVALID, "null argument");
  return run_host_pipeline(p, h_pcm, pcm_len, nullptr, nullptr, h_out, 0);
}

// ---- feature sink / ingest -------------------------------------------------------------------------------
int vadb200_scale_rows(vadb200_handle* h, float* d_rows, int64_t n_rows, double* h_stats, void* stream) {
  if (!h || n_rows < 0) return fail(VADB200_E_INVALID, "bad argument");
  if (n_rows == 0) return 0;
  if (!d_rows) return fail(VADB200_E_INVALID, "d_rows is null");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* acc = nullptr;
  CU(cudaMallocAsync(&acc, 6 * sizeof(double), st));  // stream-ordered scratch: no per-handle state
  CU(cudaMemsetAsync(acc, 0, 6 * sizeof(double), st));
  const long long total = n_rows * kNFeat;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 8ll * h->num_sms));
  for (int pass = 0; pass < 3; ++pass) {
    scale_rows_kernel<<<grid, 256, 0, st>>>(d_rows, n_rows, pass, acc);
    g_launches.fetch_add(1);
  }
  CU(cudaGetLastError());
  if (h_stats) {
    double raw[6];
    CU(cudaMemcpyAsync(raw, acc, sizeof(raw), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const double cnt = static_cast<double>(n_rows) * kNCep;
    for (int g = 0; g < 3; ++g) {
      h_stats[g] = raw[g] / cnt;
      h_stats[3 + g] = std::sqrt(raw[3 + g] / cnt);
    }
  }
  CU(cudaFreeAsync(acc, st));
  return 0;
}

int vadb200_ingest_pcm(vadb200_handle* h, const void* d_raw, int big_endian, const int64_t* d_src_start,
                       const int64_t* d_dst_start, const int64_t* d_len, int n_seg, int64_t max_len, int16_t* d_dst,
                       void* stream) {
  if (!h || n_seg < 0 || max_len < 0) return fail(VADB200_E_INVALID, "bad argument");
  if (n_seg == 0 || max_len == 0) return 0;
  if (!d_raw || !d_src_start || !d_dst_start || !d_len || !d_dst) return fail(VADB200_E_INVALID, "null argument");
  if (reinterpret_cast<uintptr_t>(d_raw) & 1) return fail(VADB200_E_INVALID, "raw sample stream must be 2-byte aligned");
  if (n_seg > 65535) return fail(VADB200_E_INVALID, "at most 65535 segments per call");
  CU(cudaSetDevice(h->device));
  const unsigned gx = static_cast<unsigned>(std::min<long long>((max_len + 255) / 256, 4ll * h->num_sms));
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t layout");
  ingest_kernel<<<dim3(gx, static_cast<unsigned>(n_seg)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(d_raw), big_endian, reinterpret_cast<const long long*>(d_src_start),
      reinterpret_cast<const long long*>(d_dst_start), reinterpret_cast<const long long*>(d_len), d_dst);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

// ---- per-frame API ---------------------------------------------------------------------------------------
static int frames_common(vadb200_handle* h, const float* d_frames, int64_t n, int frame_len, int what, float* d_out,
                         void* stream) {
  if (!h) return fail(VADB200_E_INVALID, "handle is null");
  if (n < 0 || frame_len < 1 || frame_len > kFftN) return fail(VADB200_E_INVALID, "need 1 <= frame_len <= 512, n >= 0");
  if (n == 0) return 0;
  if (!d_frames || !d_out) return fail(VADB200_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  int rc = ensure_attrs(h);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = ensure_tables(h);
  if (rc) return rc;
  const unsigned grid = static_cast<unsigned>((n + kStepFrames - 1) / kStepFrames);
  frames_kernel<<<grid, kThreads, kFramesSmemBytes, st>>>(d_frames, n, frame_len, what, d_out, h->d_tw, h->d_tw + 256);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

int vadb200_spec_frames(vadb200_handle* h, const float* d_frames, int64_t n, int frame_len, float* d_spec, void* stream) {
  return frames_common(h, d_frames, n, frame_len, 0, d_spec, stream);
}
int vadb200_mfcc_frames(vadb200_handle* h, const float* d_frames, int64_t n, int frame_len, float* d_mfcc, void* stream) {
  return frames_common(h, d_frames, n, frame_len, 1, d_mfcc, stream);
}

int vadb200_mfcc_from_spec(vadb200_handle* h, const float* d_spec, int64_t n, float* d_mfcc, void* stream) {
  if (!h) return fail(VADB200_E_INVALID, "handle is null");
  if (n < 0) return fail(VADB200_E_INVALID, "n < 0");
  if (n == 0) return 0;
  if (!d_spec || !d_mfcc) return fail(VADB200_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_tables(h);
  if (rc) return rc;
  const unsigned grid = static_cast<unsigned>((n + kStepFrames - 1) / kStepFrames);
  spec_to_mfcc_kernel<<<grid, kThreads, 0, st>>>(d_spec, n, d_mfcc);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

int vadb200_vad_windows(vadb200_handle* h, const float* d_win, int64_t n, int feat_mode, uint8_t* d_labels,
                        float* d_logits, float* d_feats, void* stream) {
  if (!h) return fail(VADB200_E_INVALID, "handle is null");
  if (!h->have_ffn) return fail(VADB200_E_STATE, "FFN weights not set (vadb200_set_ffn_weights)");
  if (n < 0) return fail(VADB200_E_INVALID, "n < 0");
  if (n == 0) return 0;
  if (!d_win) return fail(VADB200_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_tables(h);
  if (rc) return rc;
  windows_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(d_win, n, feat_mode, d_labels, d_logits, d_feats,
                                                                       h->par);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

int vadb200_ffn_predict(vadb200_handle* h, const float* d_x, int64_t n, uint8_t* d_labels, float* d_logits, void* stream) {
  if (!h) return fail(VADB200_E_INVALID, "handle is null");
  if (!h->have_ffn) return fail(VADB200_E_STATE, "FFN weights not set (vadb200_set_ffn_weights)");
  if (n < 0) return fail(VADB200_E_INVALID, "n < 0");
  if (n == 0) return 0;
  if (!d_x) return fail(VADB200_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_tables(h);
  if (rc) return rc;
  if (h->ffn_impl >= 2 && h->tc16_ok) {
    ffn_tc_rows_kernel<2><<<static_cast<unsigned>((n + 127) / 128), 128, kFfnTcSmemBytes, st>>>(
        d_x, n, h->d_tc16_blob, d_labels, d_logits, h->bias);
  } else if (h->ffn_impl >= 1) {
    ffn_tc_rows_kernel<1><<<static_cast<unsigned>((n + 127) / 128), 128, kFfnTcSmemBytes, st>>>(
        d_x, n, h->d_tc_blob, d_labels, d_logits, h->bias);
  } else {
    ffn_rows_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(d_x, n, d_labels, d_logits, h->par);
  }
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

int vadb200_get_deltas(vadb200_handle* h, const float* d_a, const float* d_b, int64_t n, float* d_out, void* stream) {
  if (!h || n < 0) return fail(VADB200_E_INVALID, "bad argument");
  if (n == 0) return 0;
  if (!d_a || !d_b || !d_out) return fail(VADB200_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  elementwise_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(0, d_a, d_b, n, 1, 0, d_out);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

int vadb200_lifter(vadb200_handle* h, const float* d_in, int64_t n_rows, int ncoef, int L, float* d_out, void* stream) {
  if (!h || n_rows < 0 || ncoef < 1) return fail(VADB200_E_INVALID, "bad argument");
  if (n_rows == 0) return 0;
  if (!d_in || !d_out) return fail(VADB200_E_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  const long long n = n_rows * ncoef;
  elementwise_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(1, d_in, nullptr, n, ncoef, L, d_out);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

// ---- streaming -------------------------------------------------------------------------------------------
int vadb200_stream_bank_create(vadb200_handle* h, int n_streams, vadb200_bank** out) {
  if (!h || !out || n_streams < 1) return fail(VADB200_E_INVALID, "bad argument");
  *out = nullptr;
  CU(cudaSetDevice(h->device));
  vadb200_bank* b = new (std::nothrow) vadb200_bank();
  if (!b) return fail(VADB200_E_NOMEM, "host allocation failed");
  b->h = h; b->n = n_streams;
  cudaError_t e = cudaMalloc(&b->d_hist, static_cast<size_t>(n_streams) * 320 * sizeof(int16_t));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_ring, static_cast<size_t>(n_streams) * 5 * kNCep * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_fed, static_cast<size_t>(n_streams) * sizeof(int));
  if (e != cudaSuccess) {
    cudaFree(b->d_hist); cudaFree(b->d_ring); cudaFree(b->d_fed);
    delete b;
    return cuda_fail(e, "vadb200_stream_bank_create");
  }
  *out = b;
  int rc = vadb200_stream_bank_reset(b, nullptr);
  if (rc) return rc;
  CU(cudaDeviceSynchronize());  // state is zero before any stream can feed
  return 0;
}

int vadb200_stream_bank_destroy(vadb200_bank* b) {
  if (!b) return 0;
  cudaSetDevice(b->h->device);
  cudaFree(b->d_hist); cudaFree(b->d_ring); cudaFree(b->d_fed);
  delete b;
  return 0;
}

int vadb200_stream_bank_reset(vadb200_bank* b, void* stream) {
  if (!b) return fail(VADB200_E_INVALID, "bank is null");
  CU(cudaSetDevice(b->h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaMemsetAsync(b->d_hist, 0, static_cast<size_t>(b->n) * 320 * sizeof(int16_t), st));
  CU(cudaMemsetAsync(b->d_ring, 0, static_cast<size_t>(b->n) * 5 * kNCep * sizeof(float), st));
  CU(cudaMemsetAsync(b->d_fed, 0, static_cast<size_t>(b->n) * sizeof(int), st));
  return 0;
}

int vadb200_stream_feed(vadb200_bank* b, const int16_t* d_chunks, uint8_t* d_labels, float* d_logits, void* stream) {
  if (!b || !d_chunks || !d_labels) return fail(VADB200_E_INVALID, "null argument");
  vadb200_handle* h = b->h;
  if (!h->have_ffn) return fail(VADB200_E_STATE, "FFN weights not set (vadb200_set_ffn_weights)");
  if (reinterpret_cast<uintptr_t>(d_chunks) & 3) return fail(VADB200_E_INVALID, "chunks must be 4-byte aligned");
  CU(cudaSetDevice(h->device));
  int rc = ensure_attrs(h);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = ensure_tables(h);
  if (rc) return rc;
  BankParams bp;
  bp.n_streams = b->n; bp.hist = b->d_hist; bp.ring = b->d_ring; bp.fed = b->d_fed;
  bp.chunks = d_chunks; bp.labels = d_labels; bp.logits = d_logits;
  bp.tw1 = h->d_tw; bp.tw2 = h->d_tw + 256; bp.feat_mode = VADB200_FEAT_ANALYSER;
  stream_feed_kernel<<<(b->n + kStepFrames - 1) / kStepFrames, kThreads, kStreamSmemBytes, st>>>(bp, h->par);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

#if defined(VADB_DEBUG_HOOKS)
// Occupancy experiment (hook build only, not in the public header): FFT phase alone at minb CTAs per SM.
int vadb200_exp_fft(vadb200_plan* p, const int16_t* d_pcm, int64_t pcm_len, int minb, void* stream) {
  vadb200_handle* h = p->h;
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FusedParams fp = base_params(p);
  fp.pcm = d_pcm; fp.pcm_len = pcm_len; fp.seg_begin = 0; fp.seg_end = static_cast<int>(p->segs.size());
  fp.counter = p->d_counter + 2 * (p->launch_seq.fetch_add(1) % kPlanCounters);
  const int grid = std::min<int>(fp.seg_end, minb * h->num_sms);
#define VADB_EXP(M)                                                                                          \
  {                                                                                                          \
    CU(cudaFuncSetAttribute(exp_fft_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, kExpSmemBytes)); \
    exp_fft_kernel<M><<<grid, kThreads, kExpSmemBytes, st>>>(fp, h->d_sink);                                 \
  }
  if (minb == 1) VADB_EXP(1) else if (minb == 2) VADB_EXP(2) else if (minb == 3) VADB_EXP(3) else VADB_EXP(4)
#undef VADB_EXP
  CU(cudaGetLastError());
  return 0;
}
#endif

// ---- FFN training (learning/ffn_trainer.py:104-175) --------------------------------------------------------
int vadb200_trainer_create(vadb200_handle* h, int64_t max_batch, float lr, float rho, float eps, vadb200_trainer** out) {
  if (!h || !out || max_batch < 1) return fail(VADB200_E_INVALID, "bad argument");
  *out = nullptr;
  CU(cudaSetDevice(h->device));
  vadb200_trainer* t = new (std::nothrow) vadb200_trainer();
  if (!t) return fail(VADB200_E_NOMEM, "host allocation failed");
  t->h = h; t->max_batch = max_batch; t->lr = lr; t->rho = rho; t->eps = eps;
  const long long n_part = (max_batch + kTrainRows - 1) / kTrainRows;
  cudaError_t e = cudaMalloc(&t->d_state, 3 * kNParams * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(t->d_state, 0, 3 * kNParams * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&t->d_partial, n_part * (kNParams + 1) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&t->d_loss, sizeof(float));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(ffn_train_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrain2SmemBytes);
  if (e != cudaSuccess) {
    cudaFree(t->d_state); cudaFree(t->d_partial); cudaFree(t->d_loss);
    delete t;
    return cuda_fail(e, "vadb200_trainer_create");
  }
  if (h->have_ffn) {  // start from the handle's classifier
    static_assert(sizeof(FfnParams) >= kNParams * sizeof(float), "FfnParams is the parameter vector");
    CU(cudaMemcpy(t->d_state, &h->par, kNParams * sizeof(float), cudaMemcpyHostToDevice));
  }
  *out = t;
  return 0;
}

int vadb200_trainer_destroy(vadb200_trainer* t) {
  if (!t) return 0;
  cudaSetDevice(t->h->device);
  cudaFree(t->d_state); cudaFree(t->d_partial); cudaFree(t->d_loss);
  delete t;
  return 0;
}

int vadb200_trainer_set_weights(vadb200_trainer* t, const float* W1, const float* b1, const float* W2, const float* b2,
                                const float* W3, const float* b3, const float* W4, const float* b4, int reset_optimizer) {
  if (!t || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !W4 || !b4) return fail(VADB200_E_INVALID, "null argument");
  std::vector<float> v(kNParams);
  const float* src[8] = {W1, b1, W2, b2, W3, b3, W4, b4};
  const int off[9] = {kOffW1, kOffB1, kOffW2, kOffB2, kOffW3, kOffB3, kOffW4, kOffB4, kNParams};
  for (int i = 0; i < 8; ++i) std::memcpy(v.data() + off[i], src[i], (off[i + 1] - off[i]) * sizeof(float));
  CU(cudaSetDevice(t->h->device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(t->d_state, v.data(), kNParams * sizeof(float), cudaMemcpyHostToDevice));
  if (reset_optimizer) CU(cudaMemset(t->d_state + kNParams, 0, 2 * kNParams * sizeof(float)));
  return 0;
}

int vadb200_trainer_get_weights(vadb200_trainer* t, float* W1, float* b1, float* W2, float* b2, float* W3, float* b3,
                                float* W4, float* b4) {
  if (!t || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !W4 || !b4) return fail(VADB200_E_INVALID, "null argument");
  std::vector<float> v(kNParams);
  CU(cudaSetDevice(t->h->device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(v.data(), t->d_state, kNParams * sizeof(float), cudaMemcpyDeviceToHost));
  float* dst[8] = {W1, b1, W2, b2, W3, b3, W4, b4};
  const int off[9] = {kOffW1, kOffB1, kOffW2, kOffB2, kOffW3, kOffB3, kOffW4, kOffB4, kNParams};
  for (int i = 0; i < 8; ++i) std::memcpy(dst[i], v.data() + off[i], (off[i + 1] - off[i]) * sizeof(float));
  return 0;
}

int vadb200_train_on_batch(vadb200_trainer* t, const float* d_x, const uint8_t* d_y, int64_t n, float* h_loss,
                           void* stream) {
  if (!t || n < 1 || !d_x || !d_y) return fail(VADB200_E_INVALID, "bad argument");
  if (n > t->max_batch) return fail(VADB200_E_INVALID, "batch larger than the trainer's max_batch");
  CU(cudaSetDevice(t->h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_tiles = static_cast<int>((n + kTrainRows - 1) / kTrainRows);
  const int n_part = std::min(n_tiles, t->h->num_sms);   // persistent CTAs: one partial gradient each
  ffn_train_grad_kernel<<<n_part, kTr2Threads, kTrain2SmemBytes, st>>>(t->d_state, d_x, d_y, n,
                                                                         1.0f / static_cast<float>(n), t->d_partial);
  ffn_train_update_kernel<<<(kNParams + 1 + 255) / 256, 256, 0, st>>>(t->d_state, t->d_state + kNParams,
                                                                       t->d_state + 2 * kNParams, t->d_partial, n_part,
                                                                       t->lr, t->rho, t->eps, t->d_loss);
  g_launches.fetch_add(2);
  CU(cudaGetLastError());
  ++t->steps;
  if (h_loss) {
    CU(cudaMemcpyAsync(h_loss, t->d_loss, sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return 0;
}

// ---- bench support -----------------------------------------------------------------------------------------
int vadb200_synth_pcm(vadb200_handle* h, int16_t* d_out, int64_t n_utt, int64_t utt_samples, int64_t utt_stride,
                      uint32_t seed, int64_t first_utt, void* stream) {
  if (!h || !d_out || n_utt < 0 || utt_samples < 0 || utt_stride < utt_samples) return fail(VADB200_E_INVALID, "bad argument");
  if (n_utt == 0 || utt_samples == 0) return 0;
  CU(cudaSetDevice(h->device));
  const long long total = n_utt * utt_samples;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 32ll * h->num_sms));
  synth_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_out, n_utt, utt_samples, utt_stride, seed, first_utt);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return 0;
}

int vadb200_fp32_peak(vadb200_handle* h, int variant, int iters, double* tflops_out) {
  if (!h || !tflops_out || iters < 1 || variant < 0 || variant > 6) return fail(VADB200_E_INVALID, "bad argument");
  CU(cudaSetDevice(h->device));
  int rc = ensure_tables(h);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  const int grid = h->num_sms * 8;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0, nullptr));
    if (variant == 0) fp32_peak_kernel<0><<<grid, 256>>>(h->d_sink, iters, 1.0000001f, 1e-9f);
    else if (variant == 1) fp32_peak_kernel<1><<<grid, 256>>>(h->d_sink, iters, 1.0000001f, 1e-9f);
    else if (variant == 2) fp32x2_peak_kernel<2><<<grid, 256>>>(h->d_sink, iters, 1.0000001f, 1e-9f);
    else if (variant == 3) fp32x2_peak_kernel<3><<<grid, 256>>>(h->d_sink, iters, 1.0000001f, 1e-9f);
    else if (variant == 4) fp32mix_peak_kernel<8><<<grid, 256>>>(h->d_sink, iters, 1.0000001f, 1e-9f);
    else if (variant == 5) fp32mix_peak_kernel<4><<<grid, 256>>>(h->d_sink, iters, 1.0000001f, 1e-9f);
    else fp32mix_peak_kernel<0><<<grid, 256>>>(h->d_sink, iters, 1.0000001f, 1e-9f);
    g_launches.fetch_add(1);
    CU(cudaEventRecord(e1, nullptr));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double per_rep = variant <= 1 ? 2.0 * 16 : variant <= 3 ? 4.0 * 16
                         : variant == 4 ? 4.0 * 8 + 2.0 * 8 : variant == 5 ? 4.0 * 8 + 2.0 * 4 : 4.0 * 8;
    const double flops = per_rep * 8 * static_cast<double>(iters) * 256.0 * grid;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return 0;
}

}  // extern "C"
