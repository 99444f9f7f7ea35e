// tools/experiments/async_and_split_kernels.cuh -- NOT part of the product build.
// A snapshot written against the headers of the round-2 mid-session build (warp_fft_quad / P2Store / ring layout have
// changed since: it documents the two designs and their numbers, it is not meant to compile against today's csrc/).
//
// Two restructurings of the fused VAD kernel that were built, verified (all 129 GPU parity tests green through
// the C ABI) and measured on a B200 in round 2, and lost to the phase-synchronous fused_kernel<2,2>:
//
//   kernel                                   ms / 1000 h     note
//   fused_kernel<2,2>  (product)             214.2           2 CTAs x 8 warps per SM, synchronous fp16 tcgen05 tile
//   fused_async_kernel                       221.5           2 CTAs per SM; fp16 weight blob resident (single PCM
//                                                            buffer, 160-slot ring, 246-row P, logE aliased), one FFN
//                                                            stage per step under the next FFT phase.  Hook ablation:
//                                                            pipeline without FFN 169.5 ms (vs 160.9 for the product
//                                                            layout: the slimmer layout and its extra per-step barrier
//                                                            cost 8.6 ms), FFN stages 53 ms (vs 62 ms synchronous) --
//                                                            the MMA waits were hidden but that was all they cost.
//   fused_split_kernel<2>                    308.0           1 CTA x 16 warps per SM: warps 0-7 FFT producers, warps
//                                                            8-15 mel / DCT / FFN consumers, double-buffered P tile,
//                                                            mbarrier hand-off, 4-slot segment queue.  Eight warps per
//                                                            role cannot hide their own latencies (the FFT alone needs
//                                                            16 resident warps: 158 -> 128 ms from 1 to 2 CTAs per SM).
//
// The code below is the exact text that was compiled inside vad_kernels.cuh (it relies on that file's helpers and
// on the tc16_* functions of ffn_tc.cuh); host dispatch was `ffn_impl` 4 (async) / 2 (split) in vadb200.cu.
#pragma once

// ---- fused_async_kernel: the VAD hot path with an asynchronous tensor-core FFN -------------------------------
// Same per-step FFT / mel / DCT pipeline as fused_kernel<2, .>, but the classifier never stops the step loop:
//   * the fp16 hi/lo weight blob (23.5 KB) is resident in shared memory for the whole kernel (the layout is
//     slimmed to make room: one PCM stage buffer, 160-slot MFCC ring, 246-row P tile, logE aliased onto the
//     transpose scratch), so tcgen05.mma can run at any time;
//   * a 128-frame tile advances ONE stage at the end of each step: boundary step = finish the previous tile
//     (logits -> labels) + window features of the new tile -> TMEM + issue layer 1; the next three steps each
//     run epilogue k (TMEM -> bias, ReLU, fp16 split -> TMEM) and issue layer k + 1.  Every layer's MMA chain
//     therefore executes under the following step's FFT phase and its mbarrier is complete when polled.
// A tile stays in flight across segment boundaries; only the kernel end drains.
struct AsyncLayout {
  static constexpr int kRingSlots = 160;                 // 128-frame tile + 4 halo + one 32-frame step, rounded
  static constexpr int kRingPitch = kRingSlots + 1;
  static constexpr int kPFirst = 10;                     // kMelFirstBin
  static constexpr int kPRows = kBins - kPFirst;
  static constexpr int offPcm = 0;
  static constexpr int offExch = offPcm + kStagePad * 2;                    // ONE stage buffer
  static constexpr int offLogE = offExch;                                   // aliased: exch is idle outside the FFT phase
  static constexpr int offP = offExch + kWarps * 2 * kExchFrame * 8;
  static constexpr int offRing = offP + kPRows * kPPitch * 4;
  static constexpr int offTw1 = offRing + ((kNCep * kRingPitch * 4 + 15) & ~15);
  static constexpr int offTw2 = offTw1 + 256 * 8;
  static constexpr int offBlob = (offTw2 + 128 * 8 + 127) & ~127;
  static constexpr int offBar = offBlob + kTc16BlobBytes;
  static constexpr int offSeg = offBar + 32;
  static constexpr int offOk = offSeg + 16;
  static constexpr int bytes = offOk + 256;
};
static_assert(kMelFirstBin == AsyncLayout::kPFirst, "P tile starts at the first mel bin");
static_assert(2 * (AsyncLayout::bytes + 1024) <= 233472, "two CTAs per SM");
static_assert(kNMel * 32 * 4 <= kWarps * 2 * kExchFrame * 8, "logE fits the transpose scratch");

__global__ void __launch_bounds__(kThreads, 2) fused_async_kernel(const __grid_constant__ FusedParams p,
                                                                  const __grid_constant__ FfnBias ffn) {
  using L = AsyncLayout;
  constexpr int kRing = L::kRingSlots, kRingPitch = L::kRingPitch;
  extern __shared__ __align__(128) unsigned char smem[];
  int16_t* s_pcm = reinterpret_cast<int16_t*>(smem + L::offPcm);
  cf2* s_exch = reinterpret_cast<cf2*>(smem + L::offExch);
  float* s_logE = reinterpret_cast<float*>(smem + L::offLogE);
  float* s_P = reinterpret_cast<float*>(smem + L::offP) - L::kPFirst * kPPitch;  // rows kPFirst .. 255 exist
  float* s_ring = reinterpret_cast<float*>(smem + L::offRing);
  cf2* s_tw1 = reinterpret_cast<cf2*>(smem + L::offTw1);
  cf2* s_tw2 = reinterpret_cast<cf2*>(smem + L::offTw2);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::offBar);  // 0: pcm, 1: weights, 2: mma
  int* s_seg = reinterpret_cast<int*>(smem + L::offSeg);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::offSeg + 8);
  unsigned char* s_ok = smem + L::offOk;
  const uint32_t w_smem = smem_u32(smem + L::offBlob);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, h = lane >> 4;
  const int fr = tid & 127, hidx = tid >> 7;                 // tile role: frame (= TMEM lane) and column half
  s_tw1[tid] = p.tw1[tid];
  if (tid < 128) s_tw2[tid] = p.tw2[tid];
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_init(&s_bar[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(s_tmem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_base = *s_tmem;
  const uint32_t tl = tm_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16);
  if (tid == 0) {  // the weights land once and stay
    mbar_arrive_expect_tx(&s_bar[1], kTc16BlobBytes);
    bulk_g2s(smem + L::offBlob, p.tc_blob, kTc16BlobBytes, &s_bar[1]);
  }
  mbar_wait(&s_bar[1], 0);

  unsigned gstep = 0;       // PCM loads waited for so far: phase parity of s_bar[0]
  uint32_t mma_par = 0;
  int pend = 0;             // stage of the tile in flight: 0 none, k = layer k issued
  long long pend_row = 0;   // output row of this thread's frame in that tile
  bool pend_valid = false;
  Segment seg;

  auto issue_load = [&](int step) {
    const long long start = seg.pcm_start + static_cast<long long>(step) * (kStepFrames * kHop);
    const long long avail = p.pcm_len - start;
    const int nsmp = avail >= kStageSamples ? kStageSamples : (avail > 0 ? static_cast<int>(avail) : 0);
    const int bulk = (nsmp * 2) & ~15;
    if (tid == 0) {
      mbar_arrive_expect_tx(&s_bar[0], static_cast<uint32_t>(bulk));
      if (bulk) bulk_g2s(s_pcm, p.pcm + start, static_cast<uint32_t>(bulk), &s_bar[0]);
      // with a single stage buffer the copy can only start after this step's FFT phase, so the NEXT step's
      // samples are pulled into L2 now: its copy then pays L2, not HBM, latency
      const long long nxt = start + kStepFrames * kHop;
      if (!(VADB_DBG(p) & 256) && nxt + kStageSamples <= p.pcm_len) bulk_prefetch_l2(p.pcm + nxt, (kStageSamples * 2) & ~15);
    }
    const int tail0 = bulk >> 1;  // < 8 samples past the last whole 16-byte chunk of the buffer
    if (tid < nsmp - tail0) s_pcm[tail0 + tid] = p.pcm[start + tail0 + tid];
  };
  // one stage of the tile in flight; `sync_first`: a CTA barrier is still owed for this step (ring / logE hazards)
  auto issue = [&](auto K, auto N, uint32_t d_col, uint32_t off) {
    if (tid == 0) {
      tc_fence_after();
      tc16_issue_layer<decltype(K)::value, decltype(N)::value>(tm_base, d_col, w_smem + off, &s_bar[2]);
    }
  };
  auto wait_mma = [&]() {
    tc_mbar_wait(&s_bar[2], mma_par);
    mma_par ^= 1u;
    tc_fence_after();
  };
  auto finish_tile = [&]() {      // pend == 4: logits -> labels
    wait_mma();
    if (hidx == 0) {
      uint32_t v[8], u[8];
      tmem_ld8(tl + kTmD4, v);
      tmem_ld8(tl + kTmD4 + kTcN4, u);
      tmem_wait_ld();
      if (pend_valid) {
        float logit[kNCls];
#pragma unroll
        for (int o = 0; o < kNCls; ++o)
          logit[o] = fmaf(__uint_as_float(v[o]) + __uint_as_float(u[o]), ffn.post[3], ffn.b4[o]);
        const bool ok = s_ok[fr] != 0 && s_ok[128 + fr] != 0;
        uint8_t lab = decide(logit);
        if (!ok) {
          logit[0] = logit[1] = logit[2] = NAN;
          lab = 0;
        }
        p.labels[pend_row] = lab;
        if (p.logits) {
          p.logits[pend_row * 3 + 0] = logit[0];
          p.logits[pend_row * 3 + 1] = logit[1];
          p.logits[pend_row * 3 + 2] = logit[2];
        }
      }
    }
    tc_fence_before();   // the next tile's tcgen05.st must not overtake these loads
    pend = 0;
  };
  // epilogue of layer `pend` + issue of the next layer; contains one CTA barrier
  auto advance = [&]() {
    wait_mma();
    if (pend == 1) tc16_hidden_epilogue<kTcN1, 2>(tl, kTmD1, ffn.b1, ffn.post[0], ffn.pre[1], hidx);
    else if (pend == 2) tc16_hidden_epilogue<kTcN2, 2>(tl, kTmD2, ffn.b2, ffn.post[1], ffn.pre[2], hidx);
    else tc16_hidden_epilogue<kTcN3, 2>(tl, kTmD3, ffn.b3, ffn.post[2], ffn.pre[3], hidx);
    tc_fence_before();
    __syncthreads();
    if (pend == 1) issue(std::integral_constant<int, kTcK2>{}, std::integral_constant<int, kTcN2>{}, kTmD2, kTc16Off2);
    else if (pend == 2) issue(std::integral_constant<int, kTcK3>{}, std::integral_constant<int, kTcN3>{}, kTmD3, kTc16Off3);
    else issue(std::integral_constant<int, kTcK4>{}, std::integral_constant<int, kTcN4>{}, kTmD4, kTc16Off4);
    ++pend;
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
    int out_done = 2;
    if (nsteps > 0) issue_load(0);   // the previous segment's last FFT phase is behind a barrier: the buffer is free

    for (int s = 0; s < nsteps; ++s, ++gstep) {
      mbar_wait(&s_bar[0], gstep & 1);
      // ---- FFT phase ---------------------------------------------------------------------
      {
        f2* ex = reinterpret_cast<f2*>(s_exch) + (warp * 2 + h) * kExchFrame;
        warp_fft_quad<13, L::kPFirst>(reinterpret_cast<const uint32_t*>(s_pcm) + (warp * 4 + h) * (kHop / 2), kHop, ex,
                                      s_tw1, s_tw2, s_P, warp * 4 + 2 * h, lane);
      }
      __syncthreads();
      if (s + 1 < nsteps) issue_load(s + 1);          // single stage buffer: refill under mel / DCT / the FFN stage
      // ---- mel + log phase (logE lives in the now idle transpose scratch) -------------------
      mel_group_dispatch<kPPitch, 32>(warp, s_P + lane, s_logE + lane);
      __syncthreads();
      // ---- DCT phase -> MFCC ring -----------------------------------------------------------
      {
        const int col = (s * kStepFrames + slot_of_col(lane)) % kRing;
        if (warp + 8 < kNCep) {
          float ra, rb;
          dct_coef2<32>(s_logE + lane, warp, warp + 8, ra, rb);
          s_ring[warp * kRingPitch + col] = ra;
          s_ring[(warp + 8) * kRingPitch + col] = rb;
        } else {
          s_ring[warp * kRingPitch + col] = dct_coef<32>(s_logE + lane, warp);
        }
      }
      // ---- FFN stage of this step --------------------------------------------------------------
      const int computed = min((s + 1) * kStepFrames, n);
      const int n_valid = computed - 2 - out_done;        // centres out_done .. computed - 3
      const bool boundary = (((s + 1) % kBlockStepsTc) == 0 || s == nsteps - 1) && n_valid > 0;  // block-uniform
      if ((VADB_DBG(p) & 8) || !boundary) {
        if (!(VADB_DBG(p) & 8) && pend >= 1 && pend <= 3) advance();  // barrier inside (also orders DCT reads before the next FFT)
        else __syncthreads();
        continue;
      }
      while (pend >= 1 && pend <= 3) advance();           // only after a short last tile: catch up synchronously
      if (pend == 4) finish_tile();
      __syncthreads();                                    // ring rows of this step visible; logE reads done
      {
        const bool valid = fr < n_valid;
        const int c = out_done + (valid ? fr : 0);
        float xl[24];
        const bool ok_half = (hidx == 0) ? window_features_range<0, 7, kRing, kRingPitch, 24>(s_ring, c, p.feat_mode, xl)
                                         : window_features_range<7, 13, kRing, kRingPitch, 24>(s_ring, c, p.feat_mode, xl);
        s_ok[tid] = ok_half ? 1 : 0;
        tc16_store_a1_half(tl, hidx, xl, ffn.pre[0]);
        pend_row = seg.out_start - p.row_base + (c - 2);
        pend_valid = valid;
        if (p.feats && valid) {
          const int k0 = hidx ? 7 : 0, nk = hidx ? 6 : 7;
          for (int k = 0; k < nk; ++k)
#pragma unroll
            for (int g = 0; g < 3; ++g) p.feats[pend_row * kNFeat + g * kNCep + k0 + k] = xl[3 * k + g];
        }
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
        issue(std::integral_constant<int, kTcK1>{}, std::integral_constant<int, kTcN1>{}, kTmD1, kTc16Off1);
        pend = 1;
      }
      out_done = max(out_done, computed - 2);
    }
  }
  // drain the tile still in flight
  while (pend >= 1 && pend <= 3) advance();
  if (pend == 4) finish_tile();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm_base, kTmemCols);
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(p.counter + 1, 1) == static_cast<int>(gridDim.x) - 1) {
      p.counter[0] = 0;
      p.counter[1] = 0;
      __threadfence();
    }
  }
}

// ---- fused_split_kernel: producer / consumer warp specialisation inside one CTA per SM ------------------------
// The phase-synchronous kernels above run FFT (FMA- and transpose-bound), mel (shared-memory-load-bound), DCT and
// the classifier one after the other behind CTA barriers, so no pipe is ever busy for long.  Here 16 warps share
// one SM as two roles that never wait for each other's phases:
//   warps 0-7  (producers): TMA-staged PCM -> packed two-frame FFT -> power tile P[buf], double-buffered;
//   warps 8-15 (consumers): mel + log -> DCT -> MFCC ring -> row flush (MODE 0 / 1) or the asynchronous
//               tensor-core FFN stages (MODE 2: fp16 hi/lo weights resident in shared memory, one stage of the
//               128-frame tile in flight per step, exactly as in fused_async_kernel).
// Hand-off is by mbarriers only: P_full / P_empty per buffer (8 warp arrivals each), pcm_full (TMA bytes) /
// pcm_empty, and a 4-slot segment queue filled one segment ahead by warp 0 from the plan's work counter.
struct SplitLayout {
  static constexpr int kThreadsAll = 512, kRoleThreads = 256;
  static constexpr int kRingSlots = 160, kRingPitch = kRingSlots + 1;
  static constexpr int kPFirst = 10, kPRows = kBins - kPFirst;
  static constexpr int kPBytes = kPRows * kPPitch * 4;
  static constexpr int offPcm = 0;                                          // 2 stage buffers
  static constexpr int offExch = offPcm + 2 * kStagePad * 2;
  static constexpr int offP = offExch + kWarps * 2 * kExchFrame * 8;        // 2 power tiles
  static constexpr int offLogE = offP + 2 * kPBytes;                        // 2 logE tiles
  static constexpr int offRing = offLogE + 2 * kNMel * 32 * 4;
  static constexpr int offTw1 = offRing + ((kNCep * kRingPitch * 4 + 15) & ~15);
  static constexpr int offTw2 = offTw1 + 256 * 8;
  static constexpr int offBlob = (offTw2 + 128 * 8 + 127) & ~127;
  static constexpr int offBar = offBlob + kTc16BlobBytes;                   // 16 mbarriers
  static constexpr int offSegQ = offBar + 16 * 8;                           // 4 x Segment
  static constexpr int offMisc = offSegQ + 4 * 32;                          // tmem address
  static constexpr int offOk = offMisc + 16;
  static constexpr int bytes = offOk + 256;
};
static_assert(SplitLayout::bytes + 1024 <= 232448, "fits one CTA");
static_assert(sizeof(Segment) <= 32, "segment queue slot");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void role_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // consumers only

template <int MODE>
__global__ void __launch_bounds__(SplitLayout::kThreadsAll, 1) fused_split_kernel(
    const __grid_constant__ FusedParams p, const __grid_constant__ FfnBias ffn) {
  using L = SplitLayout;
  constexpr int kRing = L::kRingSlots, kRingPitch = L::kRingPitch;
  extern __shared__ __align__(128) unsigned char smem[];
  int16_t* s_pcm = reinterpret_cast<int16_t*>(smem + L::offPcm);
  cf2* s_exch = reinterpret_cast<cf2*>(smem + L::offExch);
  float* s_ring = reinterpret_cast<float*>(smem + L::offRing);
  cf2* s_tw1 = reinterpret_cast<cf2*>(smem + L::offTw1);
  cf2* s_tw2 = reinterpret_cast<cf2*>(smem + L::offTw2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::offBar);
  uint64_t* pcm_full = bars + 0;   // [2] TMA bytes (+ one arrive)
  uint64_t* pcm_empty = bars + 2;  // [2] 8 producer warps
  uint64_t* P_full = bars + 4;     // [2] 8 producer warps
  uint64_t* P_empty = bars + 6;    // [2] 8 consumer warps
  uint64_t* seg_full = bars + 8;   // [4] warp 0
  uint64_t* w_full = bars + 12;    // weights
  uint64_t* mma_bar = bars + 13;
  Segment* seg_q = reinterpret_cast<Segment*>(smem + L::offSegQ);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::offMisc);
  unsigned char* s_ok = smem + L::offOk;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 256) s_tw1[tid] = p.tw1[tid];
  else if (tid < 384) s_tw2[tid - 256] = p.tw2[tid - 256];
  if (tid == 0) {
    mbar_init(&pcm_full[0], 1); mbar_init(&pcm_full[1], 1);
    mbar_init(&pcm_empty[0], kWarps); mbar_init(&pcm_empty[1], kWarps);
    mbar_init(&P_full[0], kWarps); mbar_init(&P_full[1], kWarps);
    mbar_init(&P_empty[0], kWarps); mbar_init(&P_empty[1], kWarps);
    for (int i = 0; i < 4; ++i) mbar_init(&seg_full[i], 1);
    mbar_init(w_full, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  uint32_t tm_base = 0;
  if (MODE == 2) {
    if (warp == kWarps) tmem_alloc(s_tmem, kTmemCols);
    tc_fence_before();
  }
  __syncthreads();
  if (MODE == 2) {
    tc_fence_after();
    tm_base = *s_tmem;
  }

  if (warp < kWarps) {
    // =============================== producers: PCM -> FFT -> P[buf] ===================================
    const int h = lane >> 4;
    f2* ex = reinterpret_cast<f2*>(s_exch) + (warp * 2 + h) * kExchFrame;
    Segment cur, nxt;                 // warp 0 only: one-ahead segment fetch
    nxt.n_frames = -1;
    unsigned n_loads = 0;             // warp 0: PCM loads issued so far
    unsigned n_pub = 0;               // warp 0: segments published so far
    auto publish = [&]() {            // warp 0, all lanes: fetch the next segment and publish it in the queue
      Segment sg;
      int si = 0;
      if (lane == 0) si = p.seg_begin + atomicAdd(p.counter, 1);
      si = __shfl_sync(0xffffffffu, si, 0);
      if (si < p.seg_end) sg = p.segs[si];
      else { sg.pcm_start = 0; sg.out_start = 0; sg.n_frames = -1; sg.pad = 0; }
      if (lane == 0) {
        seg_q[n_pub & 3] = sg;
        mbar_arrive(&seg_full[n_pub & 3]);   // release: the slot is written
      }
      ++n_pub;
      return sg;
    };
    auto load_pcm = [&](const Segment& sg, int step) {   // warp 0, all lanes
      const int b = n_loads & 1;
      mbar_wait(&pcm_empty[b], ((n_loads >> 1) & 1) ^ 1);   // every producer warp is done with the buffer's previous fill
      const long long start = sg.pcm_start + static_cast<long long>(step) * (kStepFrames * kHop);
      const long long avail = p.pcm_len - start;
      const int nsmp = avail >= kStageSamples ? kStageSamples : (avail > 0 ? static_cast<int>(avail) : 0);
      const int bulk = (nsmp * 2) & ~15;
      int16_t* dst = s_pcm + b * kStagePad;
      const int tail0 = bulk >> 1;  // < 8 samples past the last whole 16-byte chunk of the buffer
      if (lane < nsmp - tail0) dst[tail0 + lane] = p.pcm[start + tail0 + lane];
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(&pcm_full[b], static_cast<uint32_t>(bulk));
        if (bulk) bulk_g2s(dst, p.pcm + start, static_cast<uint32_t>(bulk), &pcm_full[b]);
      }
      ++n_loads;
    };
    if (warp == 0) {
      cur = publish();
      if (cur.n_frames > 0) load_pcm(cur, 0);
    }
    unsigned gs = 0;   // steps started by this warp (P buffer and PCM buffer index)
    for (unsigned k = 0;; ++k) {
      mbar_wait(&seg_full[k & 3], (k >> 2) & 1);
      const Segment seg = seg_q[k & 3];
      if (seg.n_frames < 0) break;
      if (warp == 0) nxt = publish();                       // one ahead
      const int nsteps = (seg.n_frames + kStepFrames - 1) / kStepFrames;
      for (int s = 0; s < nsteps; ++s, ++gs) {
        if (warp == 0) {                                     // keep the PCM pipeline one step ahead
          if (s + 1 < nsteps) load_pcm(seg, s + 1);
          else if (nxt.n_frames > 0) load_pcm(nxt, 0);
        }
        const int b = gs & 1;
        mbar_wait(&pcm_full[b], (gs >> 1) & 1);
        mbar_wait(&P_empty[b], ((gs >> 1) & 1) ^ 1);         // the consumers have read this tile's previous contents
        float* P = reinterpret_cast<float*>(smem + L::offP + b * L::kPBytes) - L::kPFirst * kPPitch;
        warp_fft_quad<13, L::kPFirst>(reinterpret_cast<const uint32_t*>(s_pcm + b * kStagePad) + (warp * 4 + h) * (kHop / 2),
                                      kHop, ex, s_tw1, s_tw2, P, warp * 4 + 2 * h, lane);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&pcm_empty[b]);
          mbar_arrive(&P_full[b]);
        }
      }
    }
  } else {
    // =============================== consumers: P[buf] -> mel -> DCT -> ring -> output ====================
    const int cw = warp - kWarps, ctid = tid - L::kRoleThreads;
    const int fr = ctid & 127, hidx = ctid >> 7;              // tile role: frame (= TMEM lane) and column half
    const uint32_t tl = tm_base + (static_cast<uint32_t>(32 * (cw & 3)) << 16);
    const uint32_t w_smem = smem_u32(smem + L::offBlob);
    uint32_t mma_par = 0;
    int pend = 0;             // stage of the tile in flight: 0 none, k = layer k issued
    long long pend_row = 0;
    bool pend_valid = false;
    if (MODE == 2) {
      if (ctid == 0) {
        mbar_arrive_expect_tx(w_full, kTc16BlobBytes);
        bulk_g2s(smem + L::offBlob, p.tc_blob, kTc16BlobBytes, w_full);
      }
      mbar_wait(w_full, 0);
    }
    auto issue = [&](auto K, auto N, uint32_t d_col, uint32_t off) {
      if (ctid == 0) {
        tc_fence_after();
        tc16_issue_layer<decltype(K)::value, decltype(N)::value>(tm_base, d_col, w_smem + off, mma_bar);
      }
    };
    auto wait_mma = [&]() {
      tc_mbar_wait(mma_bar, mma_par);
      mma_par ^= 1u;
      tc_fence_after();
    };
    auto finish_tile = [&]() {      // pend == 4: logits -> labels
      wait_mma();
      if (hidx == 0) {
        uint32_t v[8], u[8];
        tmem_ld8(tl + kTmD4, v);
        tmem_ld8(tl + kTmD4 + kTcN4, u);
        tmem_wait_ld();
        if (pend_valid) {
          float logit[kNCls];
#pragma unroll
          for (int o = 0; o < kNCls; ++o)
            logit[o] = fmaf(__uint_as_float(v[o]) + __uint_as_float(u[o]), ffn.post[3], ffn.b4[o]);
          const bool ok = s_ok[fr] != 0 && s_ok[128 + fr] != 0;
          uint8_t lab = decide(logit);
          if (!ok) {
            logit[0] = logit[1] = logit[2] = NAN;
            lab = 0;
          }
          p.labels[pend_row] = lab;
          if (p.logits) {
            p.logits[pend_row * 3 + 0] = logit[0];
            p.logits[pend_row * 3 + 1] = logit[1];
            p.logits[pend_row * 3 + 2] = logit[2];
          }
        }
      }
      tc_fence_before();   // the next tile's tcgen05.st must not overtake these loads
      pend = 0;
    };
    auto advance = [&]() {          // epilogue of layer `pend` + issue of the next layer; one consumer barrier inside
      wait_mma();
      if (pend == 1) tc16_hidden_epilogue<kTcN1, 2>(tl, kTmD1, ffn.b1, ffn.post[0], ffn.pre[1], hidx);
      else if (pend == 2) tc16_hidden_epilogue<kTcN2, 2>(tl, kTmD2, ffn.b2, ffn.post[1], ffn.pre[2], hidx);
      else tc16_hidden_epilogue<kTcN3, 2>(tl, kTmD3, ffn.b3, ffn.post[2], ffn.pre[3], hidx);
      tc_fence_before();
      role_bar();
      if (pend == 1) issue(std::integral_constant<int, kTcK2>{}, std::integral_constant<int, kTcN2>{}, kTmD2, kTc16Off2);
      else if (pend == 2) issue(std::integral_constant<int, kTcK3>{}, std::integral_constant<int, kTcN3>{}, kTmD3, kTc16Off3);
      else issue(std::integral_constant<int, kTcK4>{}, std::integral_constant<int, kTcN4>{}, kTmD4, kTc16Off4);
      ++pend;
    };

    unsigned gs = 0;
    for (unsigned k = 0;; ++k) {
      mbar_wait(&seg_full[k & 3], (k >> 2) & 1);
      const Segment seg = seg_q[k & 3];
      if (seg.n_frames < 0) break;
      const int n = seg.n_frames;
      const int nsteps = (n + kStepFrames - 1) / kStepFrames;
      int out_done = (MODE == 0) ? 0 : 2;
      for (int s = 0; s < nsteps; ++s, ++gs) {
        const int b = gs & 1;
        const float* P = reinterpret_cast<const float*>(smem + L::offP + b * L::kPBytes) - L::kPFirst * kPPitch;
        float* logE = reinterpret_cast<float*>(smem + L::offLogE + b * (kNMel * 32 * 4));
        mbar_wait(&P_full[b], (gs >> 1) & 1);
        mel_group_dispatch<kPPitch, 32>(cw, P + lane, logE + lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(&P_empty[b]);             // this warp has read all it needs from the tile
        role_bar();                                          // logE complete
        {
          const int col = (s * kStepFrames + slot_of_col(lane)) % kRing;
          if (cw + 8 < kNCep) {
            float ra, rb;
            dct_coef2<32>(logE + lane, cw, cw + 8, ra, rb);
            s_ring[cw * kRingPitch + col] = ra;
            s_ring[(cw + 8) * kRingPitch + col] = rb;
          } else {
            s_ring[cw * kRingPitch + col] = dct_coef<32>(logE + lane, cw);
          }
        }
        const int computed = min((s + 1) * kStepFrames, n);
        const bool tile_step = ((s + 1) % kBlockStepsTc) == 0 || s == nsteps - 1;
        if (MODE == 0) {
          if (!tile_step) continue;
          role_bar();                                        // ring rows visible
          float* dst = p.rows + (seg.out_start - p.row_base + out_done) * kNCep;
          const int first = out_done;
          flush_flat<L::kRoleThreads>(dst, (computed - out_done) * kNCep, ctid, [&](int j, float (&v)[4], int cnt) {
            int f = j / kNCep, cf = j - f * kNCep;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < cnt) v[i] = s_ring[cf * kRingPitch + (first + f) % kRing];
              if (++cf == kNCep) { cf = 0; ++f; }
            }
          });
          out_done = computed;
          role_bar();                                        // ring slots may be overwritten from here on
        } else if (MODE == 1) {
          if (!tile_step) continue;
          role_bar();
          const int last = computed - 3;  // centres out_done .. last
          float* dst = p.rows + (seg.out_start - p.row_base + (out_done - 2)) * kNFeat;
          const int first = out_done;
          flush_flat<L::kRoleThreads>(dst, (last - out_done + 1) * kNFeat, ctid, [&](int j, float (&v)[4], int cnt) {
            int rr = j / kNFeat, col = j - rr * kNFeat;
            int grp = col / kNCep, cf = col - grp * kNCep;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < cnt) {
                const int c = first + rr;
                const float* row = s_ring + cf * kRingPitch;
                const float c2 = row[c % kRing];
                v[i] = grp == 0 ? c2
                     : grp == 1 ? row[(c + 1) % kRing] - row[(c - 1) % kRing]
                                : (row[(c + 2) % kRing] - c2) - (c2 - row[(c - 2) % kRing]);
              }
              if (++cf == kNCep) { cf = 0; if (++grp == 3) { grp = 0; ++rr; } }
            }
          });
          out_done = max(out_done, last + 1);
          role_bar();
        } else {
          const int n_valid = computed - 2 - out_done;       // centres out_done .. computed - 3
          if (!(tile_step && n_valid > 0)) {
            if (pend >= 1 && pend <= 3) advance();
            continue;
          }
          while (pend >= 1 && pend <= 3) advance();          // only after a short last tile: catch up synchronously
          if (pend == 4) finish_tile();
          role_bar();                                        // ring rows of this step visible
          const bool valid = fr < n_valid;
          const int c = out_done + (valid ? fr : 0);
          float xl[24];
          const bool ok_half = (hidx == 0) ? window_features_range<0, 7, kRing, kRingPitch, 24>(s_ring, c, p.feat_mode, xl)
                                           : window_features_range<7, 13, kRing, kRingPitch, 24>(s_ring, c, p.feat_mode, xl);
          s_ok[ctid] = ok_half ? 1 : 0;
          tc16_store_a1_half(tl, hidx, xl, ffn.pre[0]);
          pend_row = seg.out_start - p.row_base + (c - 2);
          pend_valid = valid;
          if (p.feats && valid) {
            const int k0 = hidx ? 7 : 0, nk = hidx ? 6 : 7;
            for (int kk = 0; kk < nk; ++kk)
#pragma unroll
              for (int g = 0; g < 3; ++g) p.feats[pend_row * kNFeat + g * kNCep + k0 + kk] = xl[3 * kk + g];
          }
          tmem_wait_st();
          tc_fence_before();
          role_bar();                                        // features in TMEM; ring slots free again
          issue(std::integral_constant<int, kTcK1>{}, std::integral_constant<int, kTcN1>{}, kTmD1, kTc16Off1);
          pend = 1;
          out_done = max(out_done, computed - 2);
        }
      }
    }
    if (MODE == 2) {
      while (pend >= 1 && pend <= 3) advance();
      if (pend == 4) finish_tile();
      tc_fence_before();
    }
  }
  __syncthreads();
  if (MODE == 2 && warp == kWarps) tmem_dealloc(tm_base, kTmemCols);
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(p.counter + 1, 1) == static_cast<int>(gridDim.x) - 1) {
      p.counter[0] = 0;
      p.counter[1] = 0;
      __threadfence();
    }
  }
}

