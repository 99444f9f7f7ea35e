// Host emulation of the fused kernel's thread/lane dataflow (TEST INFRASTRUCTURE).
// Compiles vad_b200/csrc/vad_core.cuh with g++ and drives it with plain loops standing in
// for threads, __shfl_sync and shared memory, following the structure of
// vad_b200/csrc/vad_kernels.cu (32-frame steps, 16 threads per frame, 288-slot MFCC ring,
// block phases every 8 steps).  Lets the CPU test-suite check index math and fp32 accuracy
// of the exact per-thread arithmetic against the float64 oracle without a GPU.
// The product never links this file.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../vad_b200/csrc/vad_core.cuh"
#include "../../vad_b200/csrc/vad_host_tables.h"

using namespace vadb;

namespace {

constexpr int kStepFrames = 32;
constexpr int kRing = 288;
constexpr int kPPitch = 34;
constexpr int kStageSamples = (kStepFrames - 1) * kHop + kFrame;  // 5360

cf2 g_tw1[256], g_tw2[128];
FfnParams g_ffn;
bool g_init = false;

struct FrameThreads {  // registers of the 16 threads of one frame
  float xr[16][16], xi[16][16];
};
struct PairThreads {   // registers of the 16 threads of one half-warp: two frames per thread (f2 lanes)
  f2 xr[16][16], xi[16][16];
};

// The fused kernel's warp_fft_quad for one half-warp: frame A at w32a, frame B `delta` words later;
// powers go to P columns col, col + 1.
void fft_pair_pcm(const uint32_t* w32a, int delta, float* P2, int col) {
  PairThreads th;
  std::vector<f2> ex(kExchFrame);
  for (int t = 0; t < 16; ++t) {
    fft_load_pcm2(w32a, delta, t, th.xr[t], th.xi[t]);
    fft_pass1<13>(th.xr[t], th.xi[t], g_tw1, t);
  }
  for (int t = 0; t < 16; ++t) exch_store_plane(ex.data(), t, th.xr[t]);
  for (int k1 = 0; k1 < 16; ++k1) exch_load_plane(ex.data(), k1, th.xr[k1]);
  for (int t = 0; t < 16; ++t) exch_store_plane(ex.data(), t, th.xi[t]);
  for (int k1 = 0; k1 < 16; ++k1) exch_load_plane(ex.data(), k1, th.xi[k1]);
  for (int k1 = 0; k1 < 16; ++k1) dft16<16>(th.xr[k1], th.xi[k1]);
  // partner exchange: build everyone's send registers first (what __shfl_sync would read)
  static f2 sr[16][16], si[16][16];
  for (int k1 = 0; k1 < 16; ++k1)
    for (int j = 8; j < 16; ++j) {
      sr[k1][j] = (k1 == 0) ? th.xr[k1][(j + 1) & 15] : th.xr[k1][j];
      si[k1][j] = (k1 == 0) ? th.xi[k1][(j + 1) & 15] : th.xi[k1][j];
    }
  for (int k1 = 0; k1 < 16; ++k1) {
    auto xch = [&](f2 /*mine*/, int j, bool imag, int partner) { return imag ? si[partner][j] : sr[partner][j]; };
    auto store = [&](int bin, f2 v) {      // the kernels' P2Store
      if (bin >= kMelFirstBin) {
        P2[p2_index(bin, col)] = v.x;
        P2[p2_index(bin, col) + 2] = v.y;
      }
    };
    fft_split_store(th.xr[k1], th.xi[k1], k1, g_tw2, xch, store);
  }
}


void fft_frame_f32(const float* fr, int frame_len, float* Pcol) {
  FrameThreads th;
  std::vector<cf2> ex(kExchFrame);
  for (int t = 0; t < 16; ++t) {
    fft_load_f32(fr, frame_len, t, th.xr[t], th.xi[t]);
    fft_pass1<16>(th.xr[t], th.xi[t], g_tw1, t);
    exch_store(ex.data(), t, th.xr[t], th.xi[t]);
  }
  for (int k1 = 0; k1 < 16; ++k1) {
    exch_load(ex.data(), k1, th.xr[k1], th.xi[k1]);
    dft16<16>(th.xr[k1], th.xi[k1]);
  }
  float sr[16][16], si[16][16];
  for (int k1 = 0; k1 < 16; ++k1)
    for (int j = 8; j < 16; ++j) {
      sr[k1][j] = (k1 == 0) ? th.xr[k1][(j + 1) & 15] : th.xr[k1][j];
      si[k1][j] = (k1 == 0) ? th.xi[k1][(j + 1) & 15] : th.xi[k1][j];
    }
  for (int k1 = 0; k1 < 16; ++k1) {
    auto xch = [&](float, int j, bool imag, int partner) { return imag ? si[partner][j] : sr[partner][j]; };
    auto store = [&](int bin, float v) { Pcol[bin * kPPitch] = v; };
    fft_split_store(th.xr[k1], th.xi[k1], k1, g_tw2, xch, store);
  }
}

}  // namespace

extern "C" {

// weights: W1,b1,W2,b2,W3,b3,W4,b4 concatenated (5219 floats) or null (zeros).
int emul_init(const float* ffn_weights) {
  MfccConfig cfg;
  std::vector<double> fb = mel_filterbank(cfg);
  std::string why;
  if (!pack_mel_weights(fb.data(), c_tab.melw, &why)) return -1;
  folded_dct(cfg, c_tab.dct);
  pack_mel_pairs(fb.data(), c_tab.melw2);
  pack_dct_pairs(c_tab.dct, c_tab.dctp);
  fft_twiddles(g_tw1, g_tw2);
  float* dst = g_ffn.W1;
  const size_t n = kNFeat * kH1 + kH1 + kH1 * kH2 + kH2 + kH2 * kH3 + kH3 + kH3 * kNCls + kNCls;
  if (ffn_weights) std::memcpy(dst, ffn_weights, n * sizeof(float));
  else std::memset(dst, 0, n * sizeof(float));
  g_init = true;
  return 0;
}

// raw |X/512|^2 spectrum of explicit float frames (get_spec_mag, mfcc.py:59-61)
int emul_spec_f32(const float* frames, int n_frames, int frame_len, float* spec /*[n][256]*/) {
  if (!g_init) return -1;
  std::vector<float> P(256 * kPPitch);
  for (int f = 0; f < n_frames; ++f) {
    fft_frame_f32(frames + static_cast<size_t>(f) * frame_len, frame_len, P.data());
    for (int k = 0; k < 256; ++k) spec[static_cast<size_t>(f) * 256 + k] = P[k * kPPitch] * 9.5367431640625e-07f;
  }
  return 0;
}

// One utterance through the kernel's segment loop.  feat_mode: 0 analyser, 1 dataset.
// Outputs (any may be null): mfcc [T][13] (T = frames_for_length), feats [T-5][39],
// logits [T-5][3], labels [T-5].
int emul_utterance(const int16_t* pcm, long long n_samples, int feat_mode, float* mfcc_out,
                   float* feats_out, float* logits_out, uint8_t* labels_out) {
  if (!g_init) return -1;
  const long long T = frames_for_length(n_samples);
  if (T <= 0) return 0;
  // the kernel computes n = T - 1 frames in VAD mode (frame T-1 never reaches the centre of
  // the ring); the emulation computes all T so the MFCC rows can be checked too.
  const int n = static_cast<int>(T);
  const int nsteps = (n + kStepFrames - 1) / kStepFrames;
  std::vector<float> P(kP2RowsAlloc * kP2Pitch), logE(kNMel * 32), ring(kNCep * kRing);
  std::vector<int16_t> stage(kStageSamples + 8);
  int out_done = 2;
  for (int s = 0; s < nsteps; ++s) {
    const long long start = static_cast<long long>(s) * kStepFrames * kHop;
    const long long avail = std::min<long long>(n_samples - start, kStageSamples);
    std::memset(stage.data(), 0x55, stage.size() * sizeof(int16_t));  // stale garbage beyond the buffer
    std::memcpy(stage.data(), pcm + start, static_cast<size_t>(avail) * sizeof(int16_t));
    // FFT phase: warp w, half-warp h -> frame slots 4w + h and 4w + h + 2, P columns 4w + 2h, + 1
    for (int w = 0; w < 8; ++w)
      for (int h = 0; h < 2; ++h) {
        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(stage.data()) + (4 * w + h) * (kHop / 2);
        fft_pair_pcm(w32, kHop, P.data(), col_of_halfwarp(w, h));
      }
    // mel + log phase: warp g = filter group, lane = P column
    for (int g = 0; g < 8; ++g)
      for (int lane = 0; lane < 32; ++lane)
        mel2_run_dispatch<kP2Pitch, 32>(g, P.data() + 2 * lane, logE.data() + lane);
    // DCT phase: warp w -> coefficients 2w and 2w + 1 (warp 7 idles)
    for (int w = 0; w < 7; ++w)
      for (int lane = 0; lane < 32; ++lane) {
        const int f = s * kStepFrames + slot_of_col(lane);
        float ra, rb;
        dct_coef2<32>(logE.data() + lane, w, ra, rb);
        ring[2 * w * kRing + f % kRing] = ra;
        if (2 * w + 1 < kNCep) ring[(2 * w + 1) * kRing + f % kRing] = rb;
      }
    const int computed = std::min((s + 1) * kStepFrames, n);
    if (mfcc_out)
      for (int f = s * kStepFrames; f < computed; ++f)
        for (int c = 0; c < kNCep; ++c) mfcc_out[static_cast<size_t>(f) * kNCep + c] = ring[c * kRing + f % kRing];
    if (((s + 1) & 7) == 0 || s == nsteps - 1) {
      // block phase: thread i -> centre c = out_done + i, needs frames c-2 .. c+2 < computed
      // (frame T-1 is never a window member: the reference's ring is not flushed)
      const int last_center = std::min(computed - 3, n - 4);
      for (int c = out_done; c <= last_center; ++c) {
        float r[5][kNCep];
        for (int d = 0; d < 5; ++d)
          for (int k = 0; k < kNCep; ++k) r[d][k] = ring[k * kRing + (c - 2 + d) % kRing];
        float x[kNFeat];
        const bool ok = window_features(r, feat_mode, x);
        float logit[kNCls];
        ffn_forward(g_ffn, x, logit);
        uint8_t lab = decide(logit);
        if (!ok) { logit[0] = logit[1] = logit[2] = NAN; lab = 0; }
        const size_t row = static_cast<size_t>(c - 2);
        if (feats_out) std::memcpy(feats_out + row * kNFeat, x, sizeof(x));
        if (logits_out) std::memcpy(logits_out + row * kNCls, logit, sizeof(logit));
        if (labels_out) labels_out[row] = lab;
      }
      out_done = std::max(out_done, last_center + 1);
    }
  }
  return 0;
}

}  // extern "C"
