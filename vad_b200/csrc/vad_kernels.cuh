// vad_kernels.cuh -- sm_100a kernels of the fused MFCC + FFN VAD path.
//
// fused_kernel<MODE>: persistent CTAs (2 per SM, 256 threads) pull *segments* (runs of
// consecutive frames of one utterance) from an atomic work counter.  Per 32-frame step:
//   1. PCM for the step (5360 int16 = 31 hops + one frame) arrives in shared memory by one
//      TMA bulk copy (cp.async.bulk + mbarrier), double-buffered one step ahead; overlapping
//      frames are re-read from shared memory, never from HBM.
//   2. FFT phase: each warp transforms 4 frames (2 rounds x 2 frames, 16 threads per frame):
//      pruned DFT16 -> twiddle -> 16x16 transpose in shared memory -> DFT16 -> partner shuffle
//      -> real-FFT split -> |X|^2 into the P tile [256 bins][32 frames].
//   3. mel + log phase: lane = frame, warp = one of 8 balanced filter groups; weights are
//      constant-bank FFMA operands; exact-zero -> eps; log2.
//   4. DCT phase: lane = frame, warp = coefficient; folded DCT-II x lifter x log10(2) matrix;
//      result goes to a 288-slot MFCC ring in shared memory (carries the +-2 frame halo
//      across steps, so no frame is ever transformed twice inside a segment).
// Every 8 steps (256 frames) the block phase runs one thread per output frame: 5-frame window
// features, Dense 39-64-32-16-3 from constant-bank operands, argmax == VOICED, 1-byte label.
//
// Reference lines restated: mfcc.py:59-78; dataset/file_processing.py:47-70,99-101;
// realtime_analysis/sklearn_analyser.py:46-82,103-107; learning/ffn_trainer.py:104-116.
#pragma once

#include <cuda_runtime.h>

#include "vad_core.cuh"
#include "ffn_tc.cuh"

namespace vadb {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kStepFrames = 32;
constexpr int kRing = 288;
constexpr int kPPitch = 34;                                    // == 2 (mod 32): conflict-free P stores
constexpr int kStageSamples = (kStepFrames - 1) * kHop + kFrame;  // 5360
constexpr int kStagePad = 5376;                                // samples; 10752 B, 128-B multiple
constexpr int kBlockSteps = 8;                                 // block phase every 256 frames

constexpr int kOffPcm = 0;
constexpr int kOffExch = kOffPcm + 2 * kStagePad * 2;                       // 21504
constexpr int kOffP = kOffExch + kWarps * 2 * kExchFrame * 8;               // + 34816
constexpr int kOffLogE = kOffP + kBins * kPPitch * 4;                       // + 34816
constexpr int kOffRing = kOffLogE + kNMel * 32 * 4;                         // + 3328
constexpr int kOffTw1 = kOffRing + kNCep * kRing * 4;                       // + 14976
constexpr int kOffTw2 = kOffTw1 + 256 * 8;
constexpr int kOffBar = kOffTw2 + 128 * 8;
constexpr int kOffSeg = kOffBar + 32;                                       // 4 mbarriers: pcm x2, weights, mma
constexpr int kFusedSmemBytes = kOffSeg + 16;                               // s_seg, s_tmem
constexpr int kBlockStepsTc = 4;                                            // tensor-core FFN: one M=128 tile
static_assert(kTcBlobBytes <= kWarps * 2 * kExchFrame * 8 + kBins * kPPitch * 4, "weight blob must fit exch + P");

struct Segment {
  long long pcm_start;  // sample index of the first frame's first sample (multiple of 8)
  long long out_start;  // first output row of the segment
  int n_frames;         // MFCC frames to compute (outputs + 4 in window modes)
  int pad;
};

// Timing-experiment hooks (phase ablation mask, clock64 stamps) are compiled in only with
// -DVADB_DEBUG_HOOKS (tools/ab.sh builds); the product build has none of their branches.
#if defined(VADB_DEBUG_HOOKS)
#define VADB_DBG(p) ((p).debug_skip)
#else
#define VADB_DBG(p) 0
#endif

struct FusedParams {
  const int16_t* pcm;
  long long pcm_len;      // one past the last readable sample index (relative to pcm)
  const Segment* segs;
  int seg_begin, seg_end;
  int* counter;
  const cf2* tw1;
  const cf2* tw2;
  uint8_t* labels;
  float* logits;
  float* feats;
  float* rows;
  long long row_base;     // subtracted from Segment::out_start (chunked host runs)
  int feat_mode;
  const unsigned char* tc_blob;  // canonical hi/lo tf32 weight blob (ffn_tc.cuh), TC variant only
  long long* dbg_ts;             // timing experiments only: clock64 stamps of CTA 0's block phases [n][16]
  int debug_skip;                // timing experiments only (env VADB200_DEBUG_SKIP): 1 FFT, 2 mel, 4 DCT, 8 block phase
};

// ---- PTX wrappers: mbarrier + TMA bulk copy (UBLKCP) --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

struct ShflXchg {
  int lane;
  __device__ __forceinline__ float operator()(float mine, int, bool, int partner) const {
    return __shfl_sync(0xffffffffu, mine, (lane & 16) | partner);
  }
};

// Both frames of a warp (lanes 0-15 / 16-31) through the FFT; P column = frame slot fi.
template <int NZ, class LOAD>
__device__ __forceinline__ void warp_fft_pair(LOAD&& load, cf2* ex, const cf2* s_tw1, const cf2* s_tw2,
                                               float* s_P, int fi, int lane) {
  const int t = lane & 15;
  float xr[16], xi[16];
  load(xr, xi);
  fft_pass1<NZ>(xr, xi, s_tw1, t);
  exch_store(ex, t, xr, xi);
  __syncwarp();
  exch_load(ex, t, xr, xi);
  __syncwarp();
  dft16<16>(xr, xi);
  fft_split_store(xr, xi, t, s_tw2, ShflXchg{lane}, [&](int bin, float v) { s_P[bin * kPPitch + fi] = v; });
}

// One output row per thread: window features -> FFN -> decision (block phase / windows API).
__device__ __forceinline__ void classify_row(const float (&r)[5][kNCep], int feat_mode, uint8_t* labels,
                                             float* logits, float* feats, long long row) {
  float x[kNFeat];
  const bool ok = window_features(r, feat_mode, x);
  float logit[kNCls];
  ffn_forward(x, logit);
  uint8_t lab = decide(logit);
  if (!ok) {
    logit[0] = logit[1] = logit[2] = NAN;
    lab = 0;
  }
  if (labels) labels[row] = lab;
  if (logits) {
    logits[row * 3 + 0] = logit[0];
    logits[row * 3 + 1] = logit[1];
    logits[row * 3 + 2] = logit[2];
  }
  if (feats) {
#pragma unroll
    for (int i = 0; i < kNFeat; ++i) feats[row * kNFeat + i] = x[i];
  }
}

// MODE 0: MFCC rows [T][13]; 1: dataset rows [T-5][39]; 2: VAD labels [T-5].
// TC (MODE 2 only): 0 = FFN on FP32 CUDA cores, 1 = FFN on tcgen05 (tf32 x3, TMEM accumulators).
template <int MODE, int TC>
__global__ void __launch_bounds__(kThreads, 2) fused_kernel(const FusedParams p) {
  static_assert(TC == 0 || MODE == 2, "tensor-core FFN only exists in VAD mode");
  constexpr int kBlk = TC ? kBlockStepsTc : kBlockSteps;
  extern __shared__ __align__(128) unsigned char smem[];
  int16_t* s_pcm = reinterpret_cast<int16_t*>(smem + kOffPcm);
  cf2* s_exch = reinterpret_cast<cf2*>(smem + kOffExch);
  float* s_P = reinterpret_cast<float*>(smem + kOffP);
  float* s_logE = reinterpret_cast<float*>(smem + kOffLogE);
  float* s_ring = reinterpret_cast<float*>(smem + kOffRing);
  cf2* s_tw1 = reinterpret_cast<cf2*>(smem + kOffTw1);
  cf2* s_tw2 = reinterpret_cast<cf2*>(smem + kOffTw2);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  int* s_seg = reinterpret_cast<int*>(smem + kOffSeg);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, h = lane >> 4, t = lane & 15;

  s_tw1[tid] = p.tw1[tid];
  if (tid < 128) s_tw2[tid] = p.tw2[tid];
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffSeg + 8);
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_init(&s_bar[2], 1);
    mbar_init(&s_bar[3], 1);
    fence_mbar_init();
  }
  uint32_t tm_base = 0, w_par = 0, mma_par = 0;
  int dbg_n = 0;
  if (TC) {
    if (warp == 0) tmem_alloc(s_tmem, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tm_base = *s_tmem;
  }

  unsigned gstep = 0;  // loads issued so far by this CTA == steps started; buffer = gstep & 1
  Segment seg;

  // (Staggering the block-phase schedule of the two CTAs sharing an SM -- per-SM arrival rank via
  // %smid -- was tried against the lockstep hypothesis and measured neutral: 232.1 vs 231.1 ms.)

  auto issue_load = [&](int step, int buf) {
    const long long start = seg.pcm_start + static_cast<long long>(step) * (kStepFrames * kHop);
    const long long avail = p.pcm_len - start;
    const int nsmp = avail >= kStageSamples ? kStageSamples : (avail > 0 ? static_cast<int>(avail) : 0);
    const int bulk = (nsmp * 2) & ~15;
    int16_t* dst = s_pcm + buf * kStagePad;
    if (tid == 0) {
      mbar_arrive_expect_tx(&s_bar[buf], static_cast<uint32_t>(bulk));
      if (bulk) bulk_g2s(dst, p.pcm + start, static_cast<uint32_t>(bulk), &s_bar[buf]);
    }
    const int tail0 = bulk >> 1;  // < 8 samples past the last whole 16-byte chunk of the buffer
    if (tid < nsmp - tail0) dst[tail0 + tid] = p.pcm[start + tail0 + tid];
  };

  for (;;) {
    __syncthreads();
    if (tid == 0) *s_seg = p.seg_begin + atomicAdd(p.counter, 1);
    __syncthreads();
    const int si = *s_seg;
    if (si >= p.seg_end) break;
    seg = p.segs[si];
    const int n = seg.n_frames;
    const int nsteps = (n + kStepFrames - 1) / kStepFrames;
    int out_done = (MODE == 0) ? 0 : 2;
    if (nsteps > 0) issue_load(0, gstep & 1);
    __syncthreads();

    for (int s = 0; s < nsteps; ++s, ++gstep) {
      const int buf = gstep & 1;
      if (s + 1 < nsteps) issue_load(s + 1, buf ^ 1);
      mbar_wait(&s_bar[buf], (gstep >> 1) & 1);

      // ---- FFT phase ---------------------------------------------------------------------
      const uint32_t* stage32 = reinterpret_cast<const uint32_t*>(s_pcm + buf * kStagePad);
      cf2* ex = s_exch + (warp * 2 + h) * kExchFrame;
#pragma unroll 1
      for (int r = 0; r < 2; ++r) {
        if (VADB_DBG(p) & 1) break;
        const int fi = warp * 4 + r * 2 + h;
        const uint32_t* w32 = stage32 + fi * (kHop / 2);
        // (keeping the 46 twiddle values in registers instead of re-reading the shared tables was
        // measured neutral for the TC variant and spilled in the FP32-FFN variant: not used)
        warp_fft_pair<13>([&](float (&xr)[16], float (&xi)[16]) { fft_load_pcm(w32, t, xr, xi); }, ex, s_tw1,
                          s_tw2, s_P, fi, lane);
      }
      __syncthreads();

      // ---- mel + log phase -----------------------------------------------------------------
      if (!(VADB_DBG(p) & 2)) mel_group_dispatch<kPPitch, 32>(warp, s_P + lane, s_logE + lane);
      const int computed = min((s + 1) * kStepFrames, n);
      const bool block_now = (((s + 1) % kBlk) == 0 || s == nsteps - 1) && !(VADB_DBG(p) & 8);
      const bool tc_now = TC && block_now && (computed - 2 - out_done) > 0;  // block-uniform
      if (tc_now) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // exch/P generic accesses before the TMA overwrite
      __syncthreads();
      if (tc_now && tid == 0 && !(VADB_DBG(p) & 16)) {  // exch + P are idle until the next FFT phase: land the weight blob during the DCT
        mbar_arrive_expect_tx(&s_bar[2], kTcBlobBytes);
        bulk_g2s(smem + kOffExch, p.tc_blob, kTcBlobBytes, &s_bar[2]);
      }

      // ---- DCT phase -> MFCC ring -----------------------------------------------------------
      if (!(VADB_DBG(p) & 4)) {
        const int col = (s * kStepFrames + lane) % kRing;
        if (warp + 8 < kNCep) {
          float ra, rb;
          dct_coef2<32>(s_logE + lane, warp, warp + 8, ra, rb);
          s_ring[warp * kRing + col] = ra;
          s_ring[(warp + 8) * kRing + col] = rb;
        } else {
          s_ring[warp * kRing + col] = dct_coef<32>(s_logE + lane, warp);
        }
      }

      if (block_now) {
        __syncthreads();
        if (MODE == 0) {
          const long long base = (seg.out_start - p.row_base + out_done) * kNCep;
          const int total = (computed - out_done) * kNCep;
          for (int j = tid; j < total; j += kThreads) {
            const int fr = j / kNCep, cf = j - fr * kNCep;
            p.rows[base + j] = s_ring[cf * kRing + (out_done + fr) % kRing];
          }
          out_done = computed;
        } else if (MODE == 1) {
          const int last = computed - 3;  // centres out_done .. last
          const long long base = (seg.out_start - p.row_base + (out_done - 2)) * kNFeat;
          const int total = (last - out_done + 1) * kNFeat;
          for (int j = tid; j < total; j += kThreads) {
            const int rr = j / kNFeat, col = j - rr * kNFeat;
            const int grp = col / kNCep, cf = col - grp * kNCep;
            const int c = out_done + rr;
            const float* row = s_ring + cf * kRing;
            const float c2 = row[c % kRing];
            float v;
            if (grp == 0) v = c2;
            else if (grp == 1) v = row[(c + 1) % kRing] - row[(c - 1) % kRing];
            else v = (row[(c + 2) % kRing] - c2) - (c2 - row[(c - 2) % kRing]);
            p.rows[base + j] = v;
          }
          out_done = max(out_done, last + 1);
        } else if (!TC) {
          const int c = out_done + tid;
          if (c <= computed - 3) {
            float r[5][kNCep];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
              const int col = (c - 2 + d) % kRing;
#pragma unroll
              for (int k = 0; k < kNCep; ++k) r[d][k] = s_ring[k * kRing + col];
            }
            classify_row(r, p.feat_mode, p.labels, p.logits, p.feats, seg.out_start - p.row_base + (c - 2));
          }
          out_done = max(out_done, computed - 2);
        } else {
          // ---- tensor-core FFN: one M = 128 tile over all 8 warps ------------------------------------
          const int n_valid = computed - 2 - out_done;  // centres out_done .. computed-3 (<= 128)
          if (n_valid > 0) {                            // block-uniform (== tc_now)
            unsigned char* wdst = smem + kOffExch;
            {
              // thread (warp % 4, lane) = frame = TMEM lane; the two threads sharing a frame (hidx = warp / 4)
              // build the features of cepstral coefficients 0-6 / 7-12 and split every layer's columns.
              const int fr = tid & 127, hidx = tid >> 7;
              long long* ts = nullptr;
#if defined(VADB_DEBUG_HOOKS)
              if (p.dbg_ts && blockIdx.x == 0 && tid == 0 && dbg_n < 64) ts = p.dbg_ts + 16 * dbg_n;
#endif
              if (ts) ts[0] = clock64();
              const bool valid = fr < n_valid;
              const int c = out_done + (valid ? fr : 0);
              const uint32_t tl = tm_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16);
              float xl[24], logit[kNCls];
              // the half's validity flag travels through shared memory (logE is idle here): ReLU's fmaxf
              // swallows NaNs, so validity cannot be read back from the logits
              bool ok_half = true;
              if (VADB_DBG(p) & 64) {
#pragma unroll
                for (int i = 0; i < 24; ++i) xl[i] = 0.5f;
              } else {
                ok_half = (hidx == 0) ? window_features_range<0, 7, kRing, 24>(s_ring, c, p.feat_mode, xl)
                                      : window_features_range<7, 13, kRing, 24>(s_ring, c, p.feat_mode, xl);
              }
              s_logE[tid] = ok_half ? 1.0f : 0.0f;
              if (ts) ts[1] = clock64();
              tc_store_a1_half(tl, hidx, xl);
              if (p.feats && valid) {
                const long long row = seg.out_start - p.row_base + (c - 2);
                const int k0 = hidx ? 7 : 0, nk = hidx ? 6 : 7;
                for (int k = 0; k < nk; ++k)
#pragma unroll
                  for (int g = 0; g < 3; ++g) p.feats[row * kNFeat + g * kNCep + k0 + k] = xl[3 * k + g];
              }
              if (!(VADB_DBG(p) & 16)) mbar_wait(&s_bar[2], w_par);  // weight blob landed (issued before the DCT phase)
              if (ts) ts[2] = clock64();
              mma_par = ffn_tc_tile<2>(logit, tm_base, warp & 3, hidx, tid == 0, smem_u32(wdst), &s_bar[3], mma_par, ts,
                                       VADB_DBG(p));
              ++dbg_n;
              if (valid && hidx == 0) {
                const bool ok = s_logE[fr] != 0.0f && s_logE[128 + fr] != 0.0f;  // ordered by the tile's bar.syncs
                uint8_t lab = decide(logit);
                if (!ok) {
                  logit[0] = logit[1] = logit[2] = NAN;
                  lab = 0;
                }
                const long long row = seg.out_start - p.row_base + (c - 2);
                p.labels[row] = lab;
                if (p.logits) {
                  p.logits[row * 3 + 0] = logit[0];
                  p.logits[row * 3 + 1] = logit[1];
                  p.logits[row * 3 + 2] = logit[2];
                }
              }
            }
            if (!(VADB_DBG(p) & 16)) w_par ^= 1u;
            __syncthreads();  // MMAs done reading the blob before the next FFT phase rewrites exch / P
          }
          out_done = max(out_done, computed - 2);
        }
      }
    }
  }
  if (TC) {
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm_base, kTmemCols);
  }
}

// Stand-alone tensor-core FFN over feature rows [n][39] (classifier.predict duck type): one CTA =
// one 128-row tile.  Same tile routine as the fused kernel's block phase.
constexpr int kFfnTcSmemBytes = kTcBlobBytes + 64;
__global__ void __launch_bounds__(128) ffn_tc_rows_kernel(const float* x, long long n, const unsigned char* blob,
                                                          uint8_t* labels, float* logits) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcBlobBytes);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kTcBlobBytes + 32);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(s_tmem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_base = *s_tmem;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], kTcBlobBytes);
    bulk_g2s(smem, blob, kTcBlobBytes, &bars[0]);
  }
  const long long i = static_cast<long long>(blockIdx.x) * 128 + tid;
  const long long src = i < n ? i : n - 1;
  float v[kNFeat], logit[kNCls];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < kNFeat; ++k) {
    v[k] = x[src * kNFeat + k];
    ok = ok && (fabsf(v[k]) <= 3.0e38f);
  }
  mbar_wait(&bars[0], 0);
  {
    const uint32_t tl = tm_base + (static_cast<uint32_t>(32 * warp) << 16);
    float h0[24], h1[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) h0[i] = h1[i] = 0.0f;
#pragma unroll
    for (int f = 0; f < kNFeat; ++f) {
      constexpr int dummy = 0;
      (void)dummy;
      const int col = tc_feat_col(f);
      if (col < 24) h0[col] = v[f];
      else h1[col - 24] = v[f];
    }
    tc_store_a1_half(tl, 0, h0);
    tc_store_a1_half(tl, 1, h1);
  }
  ffn_tc_tile<1>(logit, tm_base, warp, 0, tid == 0, smem_u32(smem), &bars[1], 0);
  if (i < n) {
    uint8_t lab = decide(logit);
    if (!ok) {
      logit[0] = logit[1] = logit[2] = NAN;
      lab = 0;
    }
    if (labels) labels[i] = lab;
    if (logits) {
      logits[i * 3 + 0] = logit[0];
      logits[i * 3 + 1] = logit[1];
      logits[i * 3 + 2] = logit[2];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm_base, kTmemCols);
}

#if defined(VADB_DEBUG_HOOKS)
// Occupancy experiment (hook build only): the fused kernel's PCM staging + FFT phase with the power
// values folded into a per-thread checksum instead of the shared P tile, so the CTA needs only the
// PCM buffers + transpose scratch (59 KB) and 3-4 CTAs fit per SM.  MINB = min blocks per SM.
constexpr int kExpSmemBytes = 2 * kStagePad * 2 + kWarps * 2 * kExchFrame * 8 + 384 * 8 + 64;
template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) exp_fft_kernel(const FusedParams p, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  int16_t* s_pcm = reinterpret_cast<int16_t*>(smem);
  cf2* s_exch = reinterpret_cast<cf2*>(smem + 2 * kStagePad * 2);
  cf2* s_tw1 = reinterpret_cast<cf2*>(smem + 2 * kStagePad * 2 + kWarps * 2 * kExchFrame * 8);
  cf2* s_tw2 = s_tw1 + 256;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tw2 + 128);
  int* s_seg = reinterpret_cast<int*>(s_bar + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, h = lane >> 4, t = lane & 15;
  s_tw1[tid] = p.tw1[tid];
  if (tid < 128) s_tw2[tid] = p.tw2[tid];
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
  }
  unsigned gstep = 0;
  float acc = 0.0f;
  Segment seg;
  auto issue_load = [&](int step, int buf) {
    const long long start = seg.pcm_start + static_cast<long long>(step) * (kStepFrames * kHop);
    const long long avail = p.pcm_len - start;
    const int nsmp = avail >= kStageSamples ? kStageSamples : (avail > 0 ? static_cast<int>(avail) : 0);
    const int bulk = (nsmp * 2) & ~15;
    if (tid == 0) {
      mbar_arrive_expect_tx(&s_bar[buf], static_cast<uint32_t>(bulk));
      if (bulk) bulk_g2s(s_pcm + buf * kStagePad, p.pcm + start, static_cast<uint32_t>(bulk), &s_bar[buf]);
    }
  };
  for (;;) {
    __syncthreads();
    if (tid == 0) *s_seg = p.seg_begin + atomicAdd(p.counter, 1);
    __syncthreads();
    const int si = *s_seg;
    if (si >= p.seg_end) break;
    seg = p.segs[si];
    const int nsteps = (seg.n_frames + kStepFrames - 1) / kStepFrames;
    issue_load(0, gstep & 1);
    __syncthreads();
    for (int s = 0; s < nsteps; ++s, ++gstep) {
      const int buf = gstep & 1;
      if (s + 1 < nsteps) issue_load(s + 1, buf ^ 1);
      mbar_wait(&s_bar[buf], (gstep >> 1) & 1);
      const uint32_t* stage32 = reinterpret_cast<const uint32_t*>(s_pcm + buf * kStagePad);
      cf2* ex = s_exch + (warp * 2 + h) * kExchFrame;
#pragma unroll 1
      for (int r = 0; r < 2; ++r) {
        const int fi = warp * 4 + r * 2 + h;
        const uint32_t* w32 = stage32 + fi * (kHop / 2);
        float xr[16], xi[16];
        fft_load_pcm(w32, t, xr, xi);
        fft_pass1<13>(xr, xi, s_tw1, t);
        exch_store(ex, t, xr, xi);
        __syncwarp();
        exch_load(ex, t, xr, xi);
        __syncwarp();
        dft16<16>(xr, xi);
        fft_split_store(xr, xi, t, s_tw2, ShflXchg{lane}, [&](int bin, float v) { acc = fmaf(v, 1e-12f * bin, acc); });
      }
      __syncthreads();  // the PCM buffer is re-filled two steps later
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}
#endif

// ---- per-frame API kernels (explicit float32 frames; mfcc.py:59-78 one frame at a time) ----------
// One CTA = 32 frames (same step structure; frames read straight from global memory).
// what: 0 -> spectrum [n][256] (get_spec_mag), 1 -> MFCC [n][13] (get_mfcc).
constexpr int kFramesSmemBytes = kWarps * 2 * kExchFrame * 8 + kBins * kPPitch * 4 + kNMel * 32 * 4 + 384 * 8;

__global__ void __launch_bounds__(kThreads) frames_kernel(const float* frames, long long n, int frame_len, int what,
                                                         float* out, const cf2* tw1, const cf2* tw2) {
  extern __shared__ __align__(128) unsigned char smem[];
  cf2* s_exch = reinterpret_cast<cf2*>(smem);
  float* s_P = reinterpret_cast<float*>(smem + kWarps * 2 * kExchFrame * 8);
  float* s_logE = s_P + kBins * kPPitch;
  cf2* s_tw1 = reinterpret_cast<cf2*>(s_logE + kNMel * 32);
  cf2* s_tw2 = s_tw1 + 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, h = lane >> 4, t = lane & 15;
  s_tw1[tid] = tw1[tid];
  if (tid < 128) s_tw2[tid] = tw2[tid];
  __syncthreads();
  const long long f0 = static_cast<long long>(blockIdx.x) * kStepFrames;
  cf2* ex = s_exch + (warp * 2 + h) * kExchFrame;
#pragma unroll 1
  for (int r = 0; r < 2; ++r) {
    const int fi = warp * 4 + r * 2 + h;
    const long long f = min(f0 + fi, n - 1);  // clamp: surplus slots recompute the last frame
    const float* fr = frames + f * frame_len;
    warp_fft_pair<16>([&](float (&xr)[16], float (&xi)[16]) { fft_load_f32(fr, frame_len, t, xr, xi); }, ex, s_tw1,
                      s_tw2, s_P, fi, lane);
  }
  __syncthreads();
  if (what == 0) {
    // |X/512|^2 = |2X|^2 * 2^-20
    for (int j = tid; j < kStepFrames * kBins; j += kThreads) {
      const int fi = j >> 8, k = j & 255;
      if (f0 + fi < n) out[(f0 + fi) * kBins + k] = s_P[k * kPPitch + fi] * 9.5367431640625e-07f;
    }
    return;
  }
  mel_group_dispatch<kPPitch, 32>(warp, s_P + lane, s_logE + lane);
  __syncthreads();
  if (f0 + lane < n) {
    out[(f0 + lane) * kNCep + warp] = dct_coef<32>(s_logE + lane, warp);
    if (warp + 8 < kNCep) out[(f0 + lane) * kNCep + warp + 8] = dct_coef<32>(s_logE + lane, warp + 8);
  }
}

// get_mfcc_from_spec (mfcc.py:72-78) for given spectra [n][256]: P = spec * 2^20 feeds the same
// mel / log / DCT code.  One CTA = 32 spectra.
__global__ void __launch_bounds__(kThreads) spec_to_mfcc_kernel(const float* spec, long long n, float* out) {
  __shared__ float s_P[kBins * kPPitch];
  __shared__ float s_logE[kNMel * 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long f0 = static_cast<long long>(blockIdx.x) * kStepFrames;
  for (int j = tid; j < kStepFrames * kBins; j += kThreads) {
    const int fi = j >> 8, k = j & 255;
    const long long f = min(f0 + fi, n - 1);
    s_P[k * kPPitch + fi] = spec[f * kBins + k] * 1048576.0f;
  }
  __syncthreads();
  mel_group_dispatch<kPPitch, 32>(warp, s_P + lane, s_logE + lane);
  __syncthreads();
  if (f0 + lane < n) {
    out[(f0 + lane) * kNCep + warp] = dct_coef<32>(s_logE + lane, warp);
    if (warp + 8 < kNCep) out[(f0 + lane) * kNCep + warp + 8] = dct_coef<32>(s_logE + lane, warp + 8);
  }
}

// 5-frame MFCC windows [n][5][13] -> features / logits / labels (one thread per window).
__global__ void __launch_bounds__(128) windows_kernel(const float* win, long long n, int feat_mode, uint8_t* labels,
                                                      float* logits, float* feats) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r[5][kNCep];
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int k = 0; k < kNCep; ++k) r[d][k] = win[(i * 5 + d) * kNCep + k];
  classify_row(r, feat_mode, labels, logits, feats, i);
}

// classifier.predict duck type (sklearn_analyser.py:71): rows [n][39] -> class {0,1}, logits.
__global__ void __launch_bounds__(128) ffn_rows_kernel(const float* x, long long n, uint8_t* labels, float* logits) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v[kNFeat];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < kNFeat; ++k) {
    v[k] = x[i * kNFeat + k];
    ok = ok && (fabsf(v[k]) <= 3.0e38f);
  }
  float logit[kNCls];
  ffn_forward(v, logit);
  uint8_t lab = decide(logit);
  if (!ok) {
    logit[0] = logit[1] = logit[2] = NAN;
    lab = 0;
  }
  if (labels) labels[i] = lab;
  if (logits) {
    logits[i * 3 + 0] = logit[0];
    logits[i * 3 + 1] = logit[1];
    logits[i * 3 + 2] = logit[2];
  }
}

// mfcc.get_deltas (mfcc.py:81-82: a - b) and mfcc.lifter (mfcc.py:85-93) for standalone callers.
// op 0: out = a - b;  op 1: out[i] = a[i] * (1 + (L/2) sin(pi (i % ncoef) / L))  (L <= 0: copy).
__global__ void elementwise_kernel(int op, const float* a, const float* b, long long n, int ncoef, int L, float* out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (op == 0) {
    out[i] = a[i] - b[i];
  } else {
    const int k = static_cast<int>(i % ncoef);
    const float lift = (L > 0) ? 1.0f + (0.5f * L) * sinpif(static_cast<float>(k) / static_cast<float>(L)) : 1.0f;
    out[i] = lift * a[i];
  }
}

// ---- streaming bank (SKLearnAnalyzer.feed_frame for many streams, one 160-sample chunk each) -----
// State per stream (struct-of-arrays, stream fastest): hist int16[320], ring float[5][13] (slot 4 =
// newest), fed int32 (chunks fed so far).  Chunk j completes frame j-2 = hist[0:320] ++ chunk[0:80].
// Reference timing (sklearn_analyser.py:46-82): feeding frame i first classifies frame i-3 from the
// ring holding frames i-5..i-1, then pushes frame i.
struct BankParams {
  int n_streams;
  int16_t* hist;   // [n_streams][320]
  float* ring;     // [5*13][n_streams]
  int* fed;        // [n_streams]
  const int16_t* chunks;  // [n_streams][160]
  uint8_t* labels;
  float* logits;
  const cf2* tw1;
  const cf2* tw2;
  int feat_mode;
};
constexpr int kStreamFramePitch = 416;  // samples per staged frame row (208 words == 16 mod 32)
constexpr int kStreamSmemBytes = kStepFrames * kStreamFramePitch * 2 + kWarps * 2 * kExchFrame * 8 +
                                 kBins * kPPitch * 4 + kNMel * 32 * 4 + kNCep * 32 * 4 + 384 * 8;

__global__ void __launch_bounds__(kThreads) stream_feed_kernel(const BankParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  int16_t* s_fr = reinterpret_cast<int16_t*>(smem);
  cf2* s_exch = reinterpret_cast<cf2*>(smem + kStepFrames * kStreamFramePitch * 2);
  float* s_P = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_exch) + kWarps * 2 * kExchFrame * 8);
  float* s_logE = s_P + kBins * kPPitch;
  float* s_mf = s_logE + kNMel * 32;
  cf2* s_tw1 = reinterpret_cast<cf2*>(s_mf + kNCep * 32);
  cf2* s_tw2 = s_tw1 + 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, h = lane >> 4, t = lane & 15;
  const int s0 = blockIdx.x * kStepFrames;
  s_tw1[tid] = p.tw1[tid];
  if (tid < 128) s_tw2[tid] = p.tw2[tid];
  // stage frames: [hist 320 | chunk 80] per stream, then roll the history in global memory
  for (int j = tid; j < kStepFrames * 200; j += kThreads) {  // 32-bit words: 160 hist + 40 chunk per stream
    const int fi = j / 200, w = j - fi * 200;
    const int st = min(s0 + fi, p.n_streams - 1);
    const uint32_t v = (w < 160) ? reinterpret_cast<const uint32_t*>(p.hist + static_cast<long long>(st) * 320)[w]
                                 : reinterpret_cast<const uint32_t*>(p.chunks + static_cast<long long>(st) * 160)[w - 160];
    reinterpret_cast<uint32_t*>(s_fr + fi * kStreamFramePitch)[w] = v;
  }
  __syncthreads();
  // new history = old hist[160:320] ++ chunk[0:160]
  for (int j = tid; j < kStepFrames * 160; j += kThreads) {
    const int fi = j / 160, w = j - fi * 160;
    const int st = s0 + fi;
    if (st < p.n_streams) {
      const uint32_t v = (w < 80) ? reinterpret_cast<const uint32_t*>(s_fr + fi * kStreamFramePitch)[80 + w]
                                  : reinterpret_cast<const uint32_t*>(p.chunks + static_cast<long long>(st) * 160)[w - 80];
      reinterpret_cast<uint32_t*>(p.hist + static_cast<long long>(st) * 320)[w] = v;
    }
  }
  cf2* ex = s_exch + (warp * 2 + h) * kExchFrame;
#pragma unroll 1
  for (int r = 0; r < 2; ++r) {
    const int fi = warp * 4 + r * 2 + h;
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(s_fr + fi * kStreamFramePitch);
    warp_fft_pair<13>([&](float (&xr)[16], float (&xi)[16]) { fft_load_pcm(w32, t, xr, xi); }, ex, s_tw1, s_tw2,
                      s_P, fi, lane);
  }
  __syncthreads();
  mel_group_dispatch<kPPitch, 32>(warp, s_P + lane, s_logE + lane);
  __syncthreads();
  s_mf[warp * 32 + lane] = dct_coef<32>(s_logE + lane, warp);
  if (warp + 8 < kNCep) s_mf[(warp + 8) * 32 + lane] = dct_coef<32>(s_logE + lane, warp + 8);
  __syncthreads();
  if (warp == 0) {
    const int st = s0 + lane;
    if (st < p.n_streams) {
      const int fed = p.fed[st];         // chunks fed before this one
      const int frame = fed - 2;         // index of the frame completed by this chunk (< 0: none yet)
      uint8_t lab = 255;
      float lg[3] = {NAN, NAN, NAN};
      float r[5][kNCep];
      const long long ns = p.n_streams;
      if (frame >= 0) {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
          for (int k = 0; k < kNCep; ++k) r[d][k] = p.ring[(d * kNCep + k) * ns + st];
        if (frame >= 5) {  // ring holds frames frame-5 .. frame-1: classify frame-3
          float x[kNFeat];
          const bool ok = window_features(r, p.feat_mode, x);
          ffn_forward(x, lg);
          lab = decide(lg);
          if (!ok) {
            lg[0] = lg[1] = lg[2] = NAN;
            lab = 0;
          }
        }
        // push: slots shift down by one, newest in slot 4
#pragma unroll
        for (int d = 0; d < 4; ++d)
#pragma unroll
          for (int k = 0; k < kNCep; ++k) p.ring[(d * kNCep + k) * ns + st] = r[d + 1][k];
#pragma unroll
        for (int k = 0; k < kNCep; ++k) p.ring[(4 * kNCep + k) * ns + st] = s_mf[k * 32 + lane];
      }
      p.fed[st] = fed + 1;
      p.labels[st] = lab;
      if (p.logits) {
        p.logits[st * 3 + 0] = lg[0];
        p.logits[st * 3 + 1] = lg[1];
        p.logits[st * 3 + 2] = lg[2];
      }
    }
  }
}

// ---- bench support ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
// bit-identical to vad_b200/synth.py:synth_utterance
__global__ void synth_kernel(int16_t* out, long long n_utt, long long utt_samples, long long utt_stride, uint32_t seed,
                             long long first_utt) {
  const long long total = n_utt * utt_samples;
  for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long u = g / utt_samples;
    const uint32_t i = static_cast<uint32_t>(g - u * utt_samples);
    const uint32_t key = mix32(mix32(seed ^ 0x9E3779B9u) + static_cast<uint32_t>(first_utt + u));
    const uint32_t a = mix32(key + i);
    const uint32_t b = mix32(a ^ 0x85EBCA6Bu);
    const int s = static_cast<int>((a & 0xFFFFu) + (a >> 16) + (b & 0xFFFFu) + (b >> 16)) - 131070;
    const uint32_t gh = mix32((key ^ 0x5BD1E995u) + (i >> 11));
    const int gain = (gh & 0x1000u) ? static_cast<int>(1500u + (gh & 0xFFFu)) : static_cast<int>(40u + (gh & 0x7Fu));
    out[u * utt_stride + i] = static_cast<int16_t>((s * gain) >> 16);
  }
}

// FP32 FMA-pipe peak: 16 independent chains per thread. variant 0: register operands;
// variant 1: constant-bank multiplicand (the form the mel / DCT / FFN stages use).
template <int VARIANT>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* sink, int iters, float m, float a) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = static_cast<float>(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (VARIANT == 0) acc[i] = fmaf(acc[i], m, a);
        else acc[i] = fmaf(acc[i], c_par.W1[rep * 16 + i], a);
      }
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) sink[0] = s;
}

}  // namespace vadb
