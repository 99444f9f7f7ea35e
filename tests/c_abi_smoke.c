/* Pure-C consumer of include/vadb200.h (TEST): no CUDA headers, no Python.  Builds a two-utterance
 * packed batch of a deterministic int16 signal in host memory, sets FFN weights, runs the fused
 * MFCC + FFN VAD through vadb200_vad_host and prints "rows speech checksum" for the caller to
 * compare.  Exit code 0 = every call returned VADB200_OK and the row counts match the framing rule. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../include/vadb200.h"

#define CHECK(x)                                                              \
  do {                                                                        \
    int rc__ = (x);                                                           \
    if (rc__ != VADB200_OK) {                                                 \
      fprintf(stderr, "%s -> %d: %s\n", #x, rc__, vadb200_last_error());      \
      return 2;                                                               \
    }                                                                         \
  } while (0)

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s; }

int main(void) {
  vadb200_handle* h = NULL;
  vadb200_plan* plan = NULL;
  CHECK(vadb200_create(NULL, 0, &h));
  /* weights: small deterministic pattern, Keras (in,out) layout */
  static float W1[39 * 64], b1[64], W2[64 * 32], b2[32], W3[32 * 16], b3[16], W4[16 * 3], b4[3];
  uint32_t s = 12345u;
  for (int i = 0; i < 39 * 64; ++i) W1[i] = ((int)(lcg(&s) >> 16) % 2001 - 1000) * 2.4e-4f;
  for (int i = 0; i < 64 * 32; ++i) W2[i] = ((int)(lcg(&s) >> 16) % 2001 - 1000) * 2.5e-4f;
  for (int i = 0; i < 32 * 16; ++i) W3[i] = ((int)(lcg(&s) >> 16) % 2001 - 1000) * 3.5e-4f;
  for (int i = 0; i < 16 * 3; ++i) W4[i] = ((int)(lcg(&s) >> 16) % 2001 - 1000) * 5.6e-4f;
  CHECK(vadb200_set_ffn_weights(h, W1, b1, W2, b2, W3, b3, W4, b4));

  const int64_t lengths[2] = {48000, 20001};
  const int64_t offsets[2] = {0, 48000};
  const int64_t total = 48000 + 20008;
  int16_t* pcm = (int16_t*)calloc((size_t)total, sizeof(int16_t));
  for (int64_t i = 0; i < 48000 + 20001; ++i) {
    const int loud = ((i >> 12) & 1);
    pcm[i] = (int16_t)(((int)(lcg(&s) >> 16) % 6001 - 3000) / (loud ? 1 : 50));
  }
  CHECK(vadb200_plan_create(h, offsets, lengths, 2, VADB200_MODE_VAD, &plan));
  const int64_t rows = vadb200_plan_total_rows(plan);
  if (rows != vadb200_outputs_for_length(48000) + vadb200_outputs_for_length(20001)) return 3;
  uint8_t* labels = (uint8_t*)malloc((size_t)rows);
  float* logits = (float*)malloc((size_t)rows * 3 * sizeof(float));
  CHECK(vadb200_vad_host(plan, pcm, total, labels, logits, VADB200_FEAT_ANALYSER));
  int64_t speech = 0;
  double checksum = 0.0;
  for (int64_t i = 0; i < rows; ++i) {
    if (labels[i] > 1) return 4;
    speech += labels[i];
    for (int c = 0; c < 3; ++c) checksum += logits[i * 3 + c];
    /* label must agree with the returned logits: speech <=> class 1 is the arg-max */
    const float* l = logits + i * 3;
    const int want = (l[1] > l[0] && l[1] >= l[2]) ? 1 : 0;
    if (l[0] == l[0] && want != labels[i]) return 5;
  }
  /* a second FFN implementation must give the same decisions except at near-ties */
  uint8_t* labels2 = (uint8_t*)malloc((size_t)rows);
  CHECK(vadb200_set_ffn_impl(h, vadb200_get_ffn_impl(h) == 0 ? 2 : 0));
  CHECK(vadb200_vad_host(plan, pcm, total, labels2, NULL, VADB200_FEAT_ANALYSER));
  int64_t diff = 0;
  for (int64_t i = 0; i < rows; ++i) diff += labels[i] != labels2[i];
  printf("%lld %lld %.6f %lld %lld\n", (long long)rows, (long long)speech, checksum, (long long)diff,
         (long long)vadb200_launch_count());
  CHECK(vadb200_plan_destroy(plan));
  CHECK(vadb200_destroy(h));
  free(pcm); free(labels); free(labels2); free(logits);
  return diff * 200 > rows ? 6 : 0;
}
