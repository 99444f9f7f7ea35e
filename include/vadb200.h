/* vadb200.h -- C ABI of the B200-native MFCC + FFN voice-activity-detection hot path.
 *
 * The reference (nameofuser1/vad) is pure Python and has no FFI of its own; its boundary is
 * its Python call surface.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference root).  The Python package `vad_b200` binds this library
 * with ctypes and re-exports the reference's function / class names.
 *
 * Conventions: every function returns 0 on success or a negative VADB200_E_* code and never
 * throws or aborts; vadb200_last_error() gives a thread-local message.  A handle is bound to
 * one CUDA device.  The library keeps NO per-handle state in device globals: FFN weights travel
 * with every launch as kernel parameters, so different handles (= different classifiers, as one
 * SKLearnAnalyzer owns one classifier, sklearn_analyser.py:21-35) may be driven from different
 * host threads concurrently; one handle's setters (vadb200_set_*) must not race with its own
 * launches.  "d_" pointers are device memory, "h_" pointers host memory; the caller owns every
 * buffer.  `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 * device-pointer calls only enqueue work, never synchronise and can be captured in CUDA graphs.
 * A plan may be in flight on up to 16 streams at once (each launch takes its own work counter).
 */
#ifndef VADB200_H_
#define VADB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VADB200_VERSION 200

#define VADB200_OK 0
#define VADB200_E_INVALID (-1)      /* bad argument (null, negative, misaligned) */
#define VADB200_E_UNSUPPORTED (-2)  /* configuration other than the reference's config.py:20-27 */
#define VADB200_E_CUDA (-3)         /* CUDA runtime error, see vadb200_last_error() */
#define VADB200_E_NOMEM (-4)
#define VADB200_E_STATE (-5)        /* e.g. VAD requested before FFN weights were set */

/* plan modes */
#define VADB200_MODE_MFCC 0     /* rows = all T frames, 13 floats   (mfcc.get_mfcc per frame) */
#define VADB200_MODE_DATASET 1  /* rows = T-5, 39 floats [c,d1,d2]  (file_processing.process_file) */
#define VADB200_MODE_VAD 2      /* rows = T-5, uint8 label          (SKLearnAnalyzer recipe + FFN) */

/* feature recipes inside the VAD kernel */
#define VADB200_FEAT_ANALYSER 0 /* sklearn_analyser.py:52-69,103-107: z-normalised centre frame */
#define VADB200_FEAT_DATASET 1  /* file_processing.py:51-66: raw c, d1, d2 */

typedef struct vadb200_handle vadb200_handle;
typedef struct vadb200_plan vadb200_plan;
typedef struct vadb200_bank vadb200_bank;
typedef struct vadb200_trainer vadb200_trainer;

/* config.py:20-27.  Only the reference values are accepted by the fused kernels. */
typedef struct vadb200_config {
  int32_t sample_rate;  /* SAMPLERATE      16000 */
  int32_t frame_size;   /* FRAME_SIZE      400   */
  int32_t frame_step;   /* FRAME_STEP      160   */
  int32_t fft_n;        /* FFT_N           512   */
  int32_t n_filters;    /* FILTERBANKS_NUM 26    */
  int32_t n_mfcc;       /* MFCC_NUM        13    */
  double low_hz;        /* LOW_HZ          300   */
  double high_hz;       /* HIGH_HZ         8000  */
  int32_t lifter_l;     /* mfcc.lifter L   22    */
  int32_t reserved;
} vadb200_config;

const char* vadb200_last_error(void);
int vadb200_version(void);
void vadb200_default_config(vadb200_config* cfg);

/* dataset/file_processing.py:99 -- `while len(data) - offset > frame_size` (strict '>'). */
int64_t vadb200_frames_for_length(int64_t n_samples);
/* dataset/file_processing.py:40-70 -- the 5-slot ring emits T-5 rows, never flushed. */
int64_t vadb200_outputs_for_length(int64_t n_samples);

/* Replaces the per-process setup of dataset_creator.py:16 / SKLearnAnalyzer.__init__
 * (sklearn_analyser.py:21-35): builds the mel filterbank (mfcc.py:39-56), the folded
 * DCT-II x lifter matrix (mfcc.py:76-78,85-90) and FFT twiddles on `device`. */
int vadb200_create(const vadb200_config* cfg, int device, vadb200_handle** out);
int vadb200_destroy(vadb200_handle* h);

/* The dense [26][256] float64 filterbank the handle uses == mfcc.get_mel_filterbanks(). */
int vadb200_get_filterbank(vadb200_handle* h, double* h_out);

/* FFN of learning/ffn_trainer.py:104-116 (Dense 39-64-32-16-3), Keras layout W:(in,out)
 * row-major, y = x.W + b.  Host pointers; replaces model.load_weights. */
int vadb200_set_ffn_weights(vadb200_handle* h, const float* W1, const float* b1, const float* W2,
                            const float* b2, const float* W3, const float* b3, const float* W4,
                            const float* b4);

/* Where the FFN contraction runs: 0 = FP32 CUDA cores (constant-bank FFMA), 1 = tcgen05 tensor cores
 * (kind::tf32, operands split hi/lo into three MMAs, fp32 accumulation in TMEM; the default).  Both
 * meet the logit tolerance; applies to vadb200_vad_packed / vadb200_vad_host / vadb200_ffn_predict. */
int vadb200_set_ffn_impl(vadb200_handle* h, int impl);
int vadb200_get_ffn_impl(vadb200_handle* h);

/* ---- ragged packed batches (replaces dataset_creator.process_files' Pool.map over files,
 * dataset_creator.py:53-65, and split_into_frames, file_processing.py:80-103) ---------------
 * Utterance u occupies samples [offsets[u], offsets[u] + lengths[u]) of one int16 buffer;
 * offsets must be ascending multiples of 8 samples (16 bytes).  Row ranges per utterance are
 * contiguous in utterance order (vadb200_plan_row_offsets). */
int vadb200_plan_create(vadb200_handle* h, const int64_t* h_offsets, const int64_t* h_lengths,
                        int64_t n_utt, int mode, vadb200_plan** out);
int vadb200_plan_destroy(vadb200_plan* p);
int64_t vadb200_plan_total_rows(const vadb200_plan* p);
int vadb200_plan_row_offsets(const vadb200_plan* p, int64_t* h_out /* n_utt + 1 */);
/* Work decomposition of a plan: utterances are cut into segments of at most
 * vadb200_plan_segment_frames() consecutive frames (chosen from the batch size: >= 16 segments per
 * persistent CTA, at most 2048 frames).  vadb200_set_plan_segment_frames() overrides the choice
 * for plans created afterwards (0 = automatic; else a multiple of 32 in [64, 2048]); results do
 * not depend on it (tests use it to pin the benchmark's one-segment-per-utterance layout). */
int64_t vadb200_plan_segment_frames(const vadb200_plan* p);
int64_t vadb200_plan_segment_count(const vadb200_plan* p);
int vadb200_set_plan_segment_frames(vadb200_handle* h, int frames);

/* mfcc.get_mfcc over every frame (MODE_MFCC: d_out [rows][13]) or process_file's rows
 * (MODE_DATASET: d_out [rows][39]).  d_pcm must be 16-byte aligned; pcm_len = samples. */
int vadb200_mfcc_packed(vadb200_plan* p, const int16_t* d_pcm, int64_t pcm_len, float* d_out,
                        void* stream);

/* Fused MFCC -> 5-frame features -> FFN -> `argmax == VOICED` (sklearn_analyser.py:46-82 with
 * the FFN as classifier).  d_labels [rows] (1 speech / 0 non-speech); d_logits [rows][3] and
 * d_feats [rows][39] are optional (NULL).  Rows with non-finite features (sigma5 == 0) get
 * label 0 and NaN logits. */
int vadb200_vad_packed(vadb200_plan* p, const int16_t* d_pcm, int64_t pcm_len, uint8_t* d_labels,
                       float* d_logits, float* d_feats, int feat_mode, void* stream);

/* End-to-end variant with HOST buffers: chunks the batch, overlaps H2D copies, the fused kernel
 * and D2H of the labels on internal streams; returns when h_labels is complete.  Pinned host
 * memory gives full PCIe bandwidth; pageable memory works but is slower. */
int vadb200_vad_host(vadb200_plan* p, const int16_t* h_pcm, int64_t pcm_len, uint8_t* h_labels,
                     float* h_logits /* nullable */, int feat_mode);
int vadb200_set_host_chunk_samples(vadb200_handle* h, int64_t samples);
/* The same host-buffer pipeline for MODE_MFCC / MODE_DATASET plans: h_out [rows][13 | 39] float32.
 * Replaces the Pool.map(process_file) step of dataset_creator.process_files
 * (dataset_creator.py:61-65) for a packed batch of decoded files. */
int vadb200_mfcc_host(vadb200_plan* p, const int16_t* h_pcm, int64_t pcm_len, float* h_out);

/* ---- feature sink and ingest (the steps either side of the path in the offline flow) --------
 * dataset/utils.py:5-32 scale_features on packed dataset rows d_rows [n_rows][39], in place:
 * one scalar mean and one population std per group {mfcc, d1, d2} over all n_rows x 13 values,
 * accumulated in float64.  h_stats (nullable) receives {mean_mfcc, mean_d1, mean_d2, std_mfcc,
 * std_d1, std_d2}; when given, the call synchronises `stream`. */
int vadb200_scale_rows(vadb200_handle* h, float* d_rows, int64_t n_rows, double* h_stats, void* stream);
/* Decode + gather 16-bit PCM on the device: segment i copies d_len[i] samples starting at sample
 * index d_src_start[i] of the raw byte stream d_raw (2-byte aligned; big_endian != 0 swaps bytes:
 * NIST SPHERE, dataset/sph.py:33-63) to d_dst[d_dst_start[i] ...].  With the .stm segment bounds
 * of dataset/stm_parser.py:5-26 as (start, len) this is split_into_frames' concatenation
 * (dataset/file_processing.py:87-94).  max_len = the largest d_len (host copy, sizes the grid). */
int vadb200_ingest_pcm(vadb200_handle* h, const void* d_raw, int big_endian, const int64_t* d_src_start,
                       const int64_t* d_dst_start, const int64_t* d_len, int n_seg, int64_t max_len,
                       int16_t* d_dst, void* stream);

/* ---- per-frame API (mfcc.py:59-78 as the reference calls it: one float frame at a time;
 * batched here over n explicit frames of frame_len <= 512 float32 samples) ------------------ */
int vadb200_spec_frames(vadb200_handle* h, const float* d_frames, int64_t n, int frame_len,
                        float* d_spec /* [n][256] */, void* stream);
int vadb200_mfcc_frames(vadb200_handle* h, const float* d_frames, int64_t n, int frame_len,
                        float* d_mfcc /* [n][13] */, void* stream);
int vadb200_mfcc_from_spec(vadb200_handle* h, const float* d_spec /* [n][256] */, int64_t n,
                           float* d_mfcc /* [n][13] */, void* stream);

/* 5-frame MFCC windows [n][5][13] (rows t-2..t+2) -> features / logits / labels: the classify
 * step of SKLearnAnalyzer.feed_frame (sklearn_analyser.py:52-71) for whole-frame callers. */
int vadb200_vad_windows(vadb200_handle* h, const float* d_windows, int64_t n, int feat_mode,
                        uint8_t* d_labels /* nullable */, float* d_logits /* nullable [n][3] */,
                        float* d_feats /* nullable [n][39] */, void* stream);
/* classifier.predict duck type (sklearn_analyser.py:71): rows [n][39] -> class in {0,1}. */
int vadb200_ffn_predict(vadb200_handle* h, const float* d_x, int64_t n, uint8_t* d_labels,
                        float* d_logits /* nullable [n][3] */, void* stream);

/* mfcc.get_deltas (mfcc.py:81-82): out = a - b over n floats. */
int vadb200_get_deltas(vadb200_handle* h, const float* d_a, const float* d_b, int64_t n,
                       float* d_out, void* stream);
/* mfcc.lifter (mfcc.py:85-93): rows [n_rows][ncoef] times 1 + (L/2) sin(pi k / L); L <= 0 copies. */
int vadb200_lifter(vadb200_handle* h, const float* d_in, int64_t n_rows, int ncoef, int L,
                   float* d_out, void* stream);

/* ---- streaming (SKLearnAnalyzer.feed_frame, sklearn_analyser.py:46-82, for many streams) ---
 * Each stream owns a 320-sample history (two hops; a frame = history + 80 new samples) and a
 * 5-row MFCC ring on the device.  One feed =
 * one 160-sample (10 ms) chunk per stream; labels_out[i] is the decision for the frame fed
 * 3 frames earlier, or 255 while the ring is filling. */
int vadb200_stream_bank_create(vadb200_handle* h, int n_streams, vadb200_bank** out);
int vadb200_stream_bank_destroy(vadb200_bank* b);
int vadb200_stream_bank_reset(vadb200_bank* b, void* stream);
int vadb200_stream_feed(vadb200_bank* b, const int16_t* d_chunks /* [n][160], device-visible */,
                        uint8_t* d_labels /* [n], device-visible */, float* d_logits /* nullable */,
                        void* stream);

/* ---- FFN training (learning/ffn_trainer.py:104-175) ---------------------------------------
 * model.compile(loss='categorical_crossentropy', optimizer='adadelta') + model.train_on_batch for the
 * 39-64-32-16-3 network: forward, backward and the Adadelta update as two hand-written kernels per step,
 * deterministic (per-CTA partial gradients summed in a fixed order).  Keras-1 defaults: lr 1.0, rho 0.95,
 * eps 1e-8.  A trainer starts from the handle's weights if set (else zeros: call _set_weights); rows
 * d_x [n][39] float32 are what the fused kernels emit (MODE_DATASET rows after vadb200_scale_rows),
 * d_y [n] class ids 0 / 1 / 2 (config.py:45-47).  h_loss (nullable) receives the batch loss before the
 * update, as Keras reports it, and makes the call synchronise `stream`. */
int vadb200_trainer_create(vadb200_handle* h, int64_t max_batch, float lr, float rho, float eps,
                           vadb200_trainer** out);
int vadb200_trainer_destroy(vadb200_trainer* t);
int vadb200_trainer_set_weights(vadb200_trainer* t, const float* W1, const float* b1, const float* W2,
                                const float* b2, const float* W3, const float* b3, const float* W4,
                                const float* b4, int reset_optimizer);
int vadb200_trainer_get_weights(vadb200_trainer* t, float* W1, float* b1, float* W2, float* b2, float* W3,
                                float* b3, float* W4, float* b4);
int vadb200_train_on_batch(vadb200_trainer* t, const float* d_x, const uint8_t* d_y, int64_t n,
                           float* h_loss, void* stream);

/* ---- bench support -------------------------------------------------------------------------
 * Counter-based integer synthetic PCM, bit-identical to vad_b200/synth.py. */
int vadb200_synth_pcm(vadb200_handle* h, int16_t* d_out, int64_t n_utt, int64_t utt_samples,
                      int64_t utt_stride, uint32_t seed, int64_t first_utt, void* stream);
/* FP32 FMA-pipe microbenchmark: the roofline denominator (MEASURED_PEAKS.json has none).
 * variant 0: FFMA register operands, 1: FFMA constant-bank operand, 2: packed FFMA2 register operands,
 * 3: packed FFMA2 with an immediate multiplicand, 4 / 5 / 6: 8 FFMA2 chains interleaved with 8 / 4 / 0
 * scalar FFMA chains (sub-pipe co-issue probe).  Returns TFLOP/s (2 flop per FMA). */
int vadb200_fp32_peak(vadb200_handle* h, int variant, int iters, double* tflops_out);
/* number of kernel launches issued by this library since load (for bench's gpu_launches) */
int64_t vadb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VADB200_H_ */
