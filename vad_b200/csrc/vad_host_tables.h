// vad_host_tables.h -- host-side (init-time) constant tables for the fused kernels.
// Restates, in double precision, the table constructors of the reference:
//   mfcc.py:5-56   mel_from_hz / hz_from_mel / convert_to_fft_bins / get_mel_filterbanks
//   mfcc.py:76-78  scipy dct(type=2, norm='ortho')[:13]  and  mfcc.py:85-90 lifter(L=22)
// These run once per handle (the reference also builds its filterbank once:
// dataset_creator.py:16, sklearn_analyser.py:29); nothing here is on the per-frame path.
#pragma once

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "vad_core.cuh"

namespace vadb {

struct MfccConfig {
  int sample_rate = 16000;
  int frame_size = 400;
  int frame_step = 160;
  int fft_n = 512;
  int n_filters = 26;
  int n_mfcc = 13;
  double low_hz = 300.0;
  double high_hz = 8000.0;
  int lifter_l = 22;
};

// mfcc.py:5-36: n_filters + 2 bin edges (floats holding integers)
inline std::vector<double> mel_bin_edges(const MfccConfig& c) {
  const double first_mel = 1125.0 * std::log(1.0 + c.low_hz / 700.0);
  const double last_mel = 1125.0 * std::log(1.0 + c.high_hz / 700.0);
  const double delta = (last_mel - first_mel) / (c.n_filters + 1);
  std::vector<double> mels;
  for (int i = 0; i < c.n_filters + 1; ++i) mels.push_back(first_mel + i * delta);
  mels.push_back(last_mel);
  std::vector<double> bins;
  for (double m : mels) {
    const double hz = 700.0 * (std::exp(m / 1125.0) - 1.0);
    bins.push_back(std::floor((c.fft_n + 1) * hz / c.sample_rate));
  }
  return bins;
}

// mfcc.py:39-56: dense [n_filters][fft_n/2], the rising branch wins at the peak.
inline std::vector<double> mel_filterbank(const MfccConfig& c) {
  const std::vector<double> b = mel_bin_edges(c);
  const int half = c.fft_n / 2;
  std::vector<double> fb(static_cast<size_t>(c.n_filters) * half, 0.0);
  for (int m = 1; m <= c.n_filters; ++m) {
    for (int k = 0; k < half; ++k) {
      if (k >= b[m - 1] && k <= b[m]) {
        fb[(m - 1) * half + k] = (k - b[m - 1] + 0.0) / (b[m] - b[m - 1] + 0.0);
      } else if (k >= b[m] && k <= b[m + 1]) {
        fb[(m - 1) * half + k] = (b[m + 1] - k + 0.0) / (b[m + 1] - b[m] + 0.0);
      }
    }
  }
  return fb;
}

// Checks that a dense filterbank has exactly the compiled-in sparsity structure and packs the
// 444 non-zeros (x 2^-20, the 1/(2*512)^2 power scale) in kMelOff order.
inline bool pack_mel_weights(const double* fb /*[26][256]*/, float* out448, std::string* why) {
  std::memset(out448, 0, 448 * sizeof(float));
  for (int m = 0; m < kNMel; ++m) {
    for (int k = 0; k < kBins; ++k) {
      const double w = fb[m * kBins + k];
      const bool inside = (k >= kMelLo[m] && k < kMelHi[m]);
      if ((w != 0.0) != inside) {
        if (why) *why = "filterbank sparsity differs from the compiled structure at filter " +
                        std::to_string(m) + " bin " + std::to_string(k);
        return false;
      }
      if (inside) out448[kMelOff[m] + k - kMelLo[m]] = static_cast<float>(std::ldexp(w, -20));
    }
  }
  return true;
}

// Pair weights of the packed mel stage: filter m, pair row q0 + q covers bins 2 (q0 + q), + 1.
inline void pack_mel_pairs(const double* fb /*[26][256]*/, float* melw2) {
  for (int m = 0; m < kNMel; ++m)
    for (int q = 0; q < mel_nq(m); ++q)
      for (int e = 0; e < 2; ++e) {
        const int bin = 2 * (mel_q0(m) + q) + e;
        const bool inside = bin >= kMelLo[m] && bin < kMelHi[m];
        melw2[2 * (mel_qoff(m) + q) + e] = inside ? static_cast<float>(std::ldexp(fb[m * kBins + bin], -20)) : 0.0f;
      }
  melw2[2 * kMelPairs] = melw2[2 * kMelPairs + 1] = 0.0f;
}
// (M[2p][n], M[2p + 1][n]) pairs for dct_coef2 (row 13 does not exist: zeros)
inline void pack_dct_pairs(const float* dct /*[13][26]*/, float (*dctp)[kNMel][2]) {
  for (int p = 0; p < 7; ++p)
    for (int n = 0; n < kNMel; ++n) {
      dctp[p][n][0] = dct[2 * p * kNMel + n];
      dctp[p][n][1] = 2 * p + 1 < kNCep ? dct[(2 * p + 1) * kNMel + n] : 0.0f;
    }
}

// M[k][n] = lifter[k] * s_k * cos(pi k (2n+1) / 52) * log10(2): mfcc = M . log2(E)
inline void folded_dct(const MfccConfig& c, float* out /*[13][26]*/) {
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < c.n_mfcc; ++k) {
    const double lift = (c.lifter_l > 0) ? 1.0 + (c.lifter_l / 2.0) * std::sin(pi * k / c.lifter_l) : 1.0;
    const double s = (k == 0) ? std::sqrt(1.0 / c.n_filters) : std::sqrt(2.0 / c.n_filters);
    for (int n = 0; n < c.n_filters; ++n) {
      const double v = lift * s * std::cos(pi * k * (2 * n + 1) / (2.0 * c.n_filters)) * std::log10(2.0);
      out[k * c.n_filters + n] = static_cast<float>(v);
    }
  }
}

// tw1[k1*16 + t] = W256^(t k1);  tw2[k2*16 + k1] = W512^(k1 + 16 k2)
inline void fft_twiddles(cf2* tw1 /*256*/, cf2* tw2 /*128*/) {
  const double pi = 3.14159265358979323846;
  for (int k1 = 0; k1 < 16; ++k1)
    for (int t = 0; t < 16; ++t) {
      const double a = -2.0 * pi * ((t * k1) % 256) / 256.0;
      tw1[k1 * 16 + t] = cf2{static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a))};
    }
  for (int k2 = 0; k2 < 8; ++k2)
    for (int k1 = 0; k1 < 16; ++k1) {
      const double a = -2.0 * pi * (k1 + 16 * k2) / 512.0;
      tw2[k2 * 16 + k1] = cf2{static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a))};
    }
}

inline bool is_reference_config(const MfccConfig& c) {
  return c.sample_rate == 16000 && c.frame_size == kFrame && c.frame_step == kHop && c.fft_n == kFftN &&
         c.n_filters == kNMel && c.n_mfcc == kNCep && c.low_hz == 300.0 && c.high_hz == 8000.0 &&
         c.lifter_l == 22;
}

// strict '>' framing rule of dataset/file_processing.py:99
inline long long frames_for_length(long long n) { return n > kFrame ? (n - kFrame - 1) / kHop + 1 : 0; }
inline long long outputs_for_length(long long n) {
  const long long t = frames_for_length(n);
  return t > 5 ? t - 5 : 0;
}

}  // namespace vadb
