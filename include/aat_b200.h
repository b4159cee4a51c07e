/*
 * aat_b200.h — C ABI of the B200-native tokenization front end.
 *
 * Drop-in boundary for the hot path of mrsndmn/audio-adaptive-tokenizer:
 * log-mel -> adaptive segment boundaries -> ragged per-segment mean-pool.
 * The reference is pure Python; the binding a maintainer adds is a ctypes stub
 * (INTEGRATION.md).  Every entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no C++/torch types; `void *stream` is a cudaStream_t
 *     (NULL = legacy default stream).
 *   - return value: 0 = AAT_OK, negative = aat_status; aat_last_error() gives a message
 *     (thread-local).  No exceptions cross the ABI.
 *   - `*_dev` pointers are device memory owned by the caller (e.g. torch tensors);
 *     `*_host` pointers are host memory.  The library never frees or reallocates
 *     caller memory.  Device entry points are asynchronous on `stream` and
 *     allocate nothing (safe inside CUDA-graph capture); `aat_host_*` entry
 *     points copy host<->device themselves and synchronise before returning.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     AAT_ERR_CUDA.
 *   - aat_logmel, aat_boundaries and aat_segment_mean_pool are launched with programmatic dependent launch:
 *     enqueued back to back on one stream, each kernel's prologue overlaps the tail of the one before it and waits
 *     on the device before touching its predecessor's outputs.  Ordering against any other work on the stream is the
 *     usual one: nothing a kernel of this library reads is touched before that wait, unless the caller opts in
 *     (AAT_POOL_EMB_READY).  A plan may have one launch of each in flight at a time (use one plan per stream): the
 *     plan owns the kernels' device-side scheduling state (tile counter, completion ticket, look-back words, the
 *     pool kernel's cross-CTA partial sums).  Plans on different streams are independent and may run concurrently.
 *     Forward progress when persistent grids of several launches share the GPU: a CTA only ever waits for CTAs of
 *     its own launch, and the hardware hands out the CTAs of a launch in index order.  The boundary kernel looks back
 *     (lower indices: resident or done).  A pool CTA publishes its pieces before it waits for anything and waits for
 *     higher indices only, so of the resident CTAs of a partly resident launch at most the last ones wait for a CTA
 *     that has no slot yet; the others leave, and their slots go to the CTAs waited for.  Partly resident launches
 *     therefore cannot block each other for good (exercised by the tests with two to four plans in flight and by
 *     every bench run with six).
 *   - device entry points run on the context's device whatever the calling thread's current device is; `stream`
 *     must belong to that device.
 *
 * Packed batch layout ("plan"): a batch of B utterances with n_samples[b] samples each.
 *   wave      : concatenated samples, utterance b at wave_off[b] = sum_{i<b} n_samples[i]
 *   mel       : utterance b is a C-contiguous (n_mels, T_b) float32 block starting at
 *               element n_mels * frame_off[b], T_b = 1 + n_samples[b] / hop
 *               (for equal lengths this is exactly a [B, n_mels, T] tensor)
 *   amp       : float32 per mel frame, utterance b at frame_off[b]
 *   segments  : utterance b owns slots [seg_slot_off[b], seg_slot_off[b+1]) of
 *               seg_start / seg_len; seg_count[b] of them are valid
 */
#ifndef AAT_B200_H
#define AAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define AAT_B200_VERSION 201 /* 0.2.1: + aat_utterance_frame_csr; 0.2.0: aat_segment_mean_pool takes a plan and flags;
                                n_seg_dev holds {S, frames} */

typedef enum aat_status {
    AAT_OK = 0,
    AAT_ERR_INVALID = -1,     /* bad argument (NULL, negative size, misaligned pointer) */
    AAT_ERR_UNSUPPORTED = -2, /* configuration outside what the kernels implement   */
    AAT_ERR_CUDA = -3,        /* CUDA runtime error / no device                      */
    AAT_ERR_CAPACITY = -4,    /* an output buffer was too small (segments, minima)   */
    AAT_ERR_TAIL = -5         /* reference would raise: tail longer than min_segment_frames
                                 (ref:src/aat/tokenizer.py:102-105 broadcast error)  */
} aat_status;

typedef enum aat_dtype {
    AAT_F32 = 0,
    AAT_F64 = 1,
    AAT_F16 = 2,
    AAT_BF16 = 3
} aat_dtype;

/* Constructor arguments of AdaptiveAudioAmplitudeTokenizer (ref:src/aat/tokenizer.py:15-38),
 * with the two derived frame counts already evaluated by the caller
 * (milliseconds_to_frames, ref:src/aat/tokenizer.py:94-95). */
typedef struct aat_config {
    int32_t sampling_rate;          /* 16000 */
    int32_t n_fft;                  /* 400 (the only length the FFT kernel implements) */
    int32_t hop_length;             /* 160; 1..n_fft */
    int32_t num_mel_filters;        /* 64; 1..128 */
    int32_t running_mean_points;    /* 12; 1..2040 */
    int32_t reserved0;
    int64_t min_segment_frames;     /* 2000 */
    int64_t max_segment_frames;     /* 24000; > 0 */
    float max_amplitude_for_minima; /* 15 */
    int32_t reserved1;
} aat_config;

typedef struct aat_ctx aat_ctx;   /* per-device constant tables + scratch; immutable after create */
typedef struct aat_plan aat_plan; /* device-resident layout tables of one batch shape           */

/* ------------------------------------------------------------------ library */
int aat_version(void);
const char *aat_last_error(void);
/* Number of CUDA kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t aat_kernel_launch_count(void);
/* A caller that replays a CUDA graph holding n of the library's kernel launches (captured through these entry points,
 * where they were counted once) reports the replay here, so that the count stays the number of kernels that RAN. */
void aat_kernel_launch_count_add(int64_t n);

/* ------------------------------------------------------------------ profiling
 * Optional CUDA-event timing of the library's own kernels, recorded on the launching stream
 * immediately before and after each launch (so a caller can attribute time to one kernel even
 * though the launch sits inside a C call).  Kernel ids: */
typedef enum aat_kernel_id {
    AAT_K_LOGMEL = 0,
    AAT_K_BOUNDARIES = 1,
    AAT_K_FRAME_CSR = 2,
    AAT_K_POOL = 3,
    AAT_K_COUNT = 4
} aat_kernel_id;
/* kernel_mask: bit i enables kernel id i; 0 disables.  Resets the counters.  Must not be enabled
 * while a stream capture is in progress.  At most 16384 launches are recorded per enable. */
int aat_profile_enable(aat_ctx *ctx, uint32_t kernel_mask);
/* Synchronises the recorded events and returns, per kernel id, the number of recorded launches and
 * the sum of their durations in milliseconds (arrays of AAT_K_COUNT entries). */
int aat_profile_summary(aat_ctx *ctx, int64_t *launches, double *total_ms);
/* Record only every `every`-th launch of each enabled kernel (default 1 = all).  An event between two kernels
 * forces the second to start after the first has drained, so sampling keeps most launches of a timed loop free to
 * overlap (programmatic dependent launch) while the sampled ones are timed in isolation.  Takes effect at the next
 * aat_profile_enable. */
int aat_profile_sample_every(aat_ctx *ctx, int32_t every);

/* ------------------------------------------------------------------ context
 * Replaces AdaptiveAudioAmplitudeTokenizer.__init__ (ref:src/aat/tokenizer.py:15-53).
 * window_host      : n_fft float64, the analysis window (window_function(n_fft,"hann"), ref :51)
 * mel_filters_host : (n_fft/2+1, num_mel_filters) float64 row-major (mel_filter_bank, ref :41-49)
 * Both tables are uploaded as given, so the device uses bit-identical constants to the
 * tokenizer attributes `window_fn` / `mel_filters`. */
int aat_create(int device, const aat_config *cfg, const double *window_host, const double *mel_filters_host,
               aat_ctx **out);
int aat_destroy(aat_ctx *ctx);
int aat_get_config(const aat_ctx *ctx, aat_config *out);

/* ------------------------------------------------------------------ plan
 * Layout tables for a batch of n_utts utterances (sizes live on the host, as array shapes
 * do in the reference).  Creation uploads a few small tables and synchronises; reuse the
 * plan for every batch of the same shape. */
int aat_plan_create(aat_ctx *ctx, int32_t n_utts, const int64_t *n_samples_host, aat_plan **out);
int aat_plan_destroy(aat_plan *plan);
int64_t aat_plan_total_samples(const aat_plan *plan);
int64_t aat_plan_total_frames(const aat_plan *plan);    /* sum of T_b                                */
int64_t aat_plan_total_seg_slots(const aat_plan *plan); /* sum of per-utterance segment capacities   */
/* Copies host-side copies of the tables: each array has n_utts+1 entries. NULL = skip. */
int aat_plan_offsets(const aat_plan *plan, int64_t *wave_off_host, int64_t *frame_off_host,
                     int64_t *seg_slot_off_host);

/* ------------------------------------------------------------------ K1+K2: log-mel
 * Replaces get_melspec -> transformers.audio_utils.spectrogram(power=2, mel_filters, "log10")
 * (ref:src/aat/tokenizer.py:107-119, TF:audio_utils.py:769-830): reflect pad n_fft/2, frames of
 * n_fft at hop, window, real DFT in float64, spectrum rounded to complex64, |.|^2 in float64,
 * mel projection, max(1e-10, .), log10, float32.
 * wave_dev   : packed samples, AAT_F32 or AAT_F64
 * znorm_stats_dev : optional float64 [2 * n_utts] (mean, population variance per utterance, as aat_normalize writes
 *              into stats_dev): the call-site normalisation (x - mean) / (std + 1e-6) in float64
 *              (ref:src/aat/training/collate.py:135-152, ref:scripts/audio_tokenization_melspec.py:40) is then applied
 *              to every sample as it is staged, bit for bit what aat_normalize(AAT_NORM_ZSCORE, float64 out) followed
 *              by this call gives, without writing or re-reading a normalised waveform.  NULL = samples as they are
 * mel_dev    : packed (n_mels, T_b) float32 blocks (see layout above)
 * amp_dev    : optional (may be NULL) float32 per frame: -10 * mean over mels of the float32
 *              log-mel, accumulated in the order numpy uses (ref:src/aat/tokenizer.py:67)
 * One launch per plan may be in flight at a time (the plan owns the kernel's tile counter). */
int aat_logmel(aat_ctx *ctx, const aat_plan *plan, const void *wave_dev, int wave_dtype, const double *znorm_stats_dev,
               float *mel_dev, float *amp_dev, void *stream);

/* The amplitude curve alone: amp_dev[frame] = -10 * mean over mels of mel_dev, in numpy's float32 order
 * (ref:src/aat/tokenizer.py:67) — what aat_logmel's fused epilogue writes, as a separate fully parallel pass over
 * a packed log-mel (the plan's, or the reference's own).  Pipelined schedules prefer it: see aat_b200/pipeline.py. */
int aat_amplitude(aat_ctx *ctx, const aat_plan *plan, const float *mel_dev, float *amp_dev, void *stream);

/* ------------------------------------------------------------------ K3: boundaries
 * Replaces find_amplitude_minimas + pretokenize + process_segments_boarders
 * (ref:src/aat/tokenizer.py:55-92, 121-139, 141-183).  Bit-exact: sequential float32
 * column mean / cumsum / running mean, strict local maxima with the 1e-5 epsilon,
 * loudness gate, x hop, append N, merge < min, split > max (np.split clamp semantics), padded tail.
 * mel_dev        : packed log-mel as written by aat_logmel (or the reference's own mel); may be
 *                  NULL when amp_dev is given
 * amp_dev        : optional precomputed amplitude curve (from aat_logmel); NULL = derive from mel_dev
 * seg_start_dev,
 * seg_len_dev    : int64, aat_plan_total_seg_slots entries; start sample and length of each segment
 * seg_count_dev  : int32 [n_utts]
 * minima_dev     : optional int64, packed per utterance at frame_off[b] (capacity T_b); mel-frame indices
 * minima_count_dev : optional int32 [n_utts]
 * status_dev     : int32 [n_utts]; negative = aat_status (AAT_ERR_CAPACITY / AAT_ERR_TAIL) for that utterance,
 *                  otherwise bit 0 is set when the last segment is the zero-padded tail
 *                  (ref:src/aat/tokenizer.py:177-181)
 * seg_off_dev, n_seg_dev, utt_seg_off_dev : optional (NULL = skip); when given, the kernel also emits the packed
 *                  frame CSR (every utterance's CTA writes its own slice after a look-back over the utterances
 *                  before it), exactly what aat_segment_frame_csr would write; n_seg_dev is int64 [2]:
 *                  {S, seg_off[S]} = segments of the batch and the HuBERT frames they cover
 * One launch per plan may be in flight at a time (the plan owns the completion ticket and the look-back words). */
int aat_boundaries(aat_ctx *ctx, const aat_plan *plan, const float *mel_dev, const float *amp_dev,
                   int64_t *seg_start_dev, int64_t *seg_len_dev, int32_t *seg_count_dev, int64_t *minima_dev,
                   int32_t *minima_count_dev, int32_t *status_dev, int64_t *seg_off_dev, int64_t *n_seg_dev,
                   int64_t *utt_seg_off_dev, void *stream);

/* Replaces process_segments_boarders alone (ref:src/aat/tokenizer.py:141-183) for one utterance:
 * boarders_dev [n_boarders] int64 -> (seg_start, seg_len)[capacity], *seg_count_dev, *status_dev. */
int aat_process_boarders(aat_ctx *ctx, int64_t n_samples, const int64_t *boarders_dev, int64_t n_boarders,
                         int64_t *seg_start_dev, int64_t *seg_len_dev, int64_t capacity, int32_t *seg_count_dev,
                         int32_t *status_dev, void *stream);

/* Segment lengths (samples) -> CSR offsets in HuBERT-frame units for the pool kernel, on device.
 * Per-segment encode convention consumed by ref:scripts/mean_hubert_embeddings.py:18-20:
 * n_i = max(0, (L_i - 400) / 320 + 1) frames (TF:models/hubert/modeling_hubert.py:675-688).
 * seg_off_dev   : int64 [total_seg_slots + 1]; entries [0, S] are written
 * n_seg_dev     : int64 [2]; {S, seg_off[S]}: total number of segments in the batch and the frames they cover
 * utt_seg_off_dev : optional int64 [n_utts+1]; first packed segment index of each utterance */
int aat_segment_frame_csr(aat_ctx *ctx, const aat_plan *plan, const int64_t *seg_len_dev,
                          const int32_t *seg_count_dev, int64_t *seg_off_dev, int64_t *n_seg_dev,
                          int64_t *utt_seg_off_dev, void *stream);

/* The same CSR for the WHOLE-UTTERANCE encode convention (convention (ii)): the encoder ran once
 * over every utterance and emb_dev holds the T_b = max(0, (N_b - 400) / 320 + 1) rows of utterance b back to back (plan
 * order).  The segment that starts at sample s of utterance b starts at row min(s / 320, T_b) of that utterance — the
 * integer division by the encoder's stride that the collator applies to mel frames with `// hop_length`
 * (ref:src/aat/training/collate.py:340) — and the last segment of an utterance ends at T_b (the lengths add up to at
 * least N_b, ref:src/aat/tokenizer.py:195), so every encoder row belongs to exactly one segment.  A padded tail that
 * starts behind the last encoder row is an empty segment (NaN mean, like torch's mean of no rows).
 * seg_start_dev / seg_count_dev / utt_seg_off_dev : as written by aat_boundaries (utt_seg_off_dev also by
 *                 aat_segment_frame_csr)
 * seg_off_dev   : int64 [total_seg_slots + 1]; entries [0, S] are written — a buffer of its own if the per-segment CSR
 *                 of the same batch is still needed
 * n_seg_dev     : int64 [2]; {S, sum of T_b}
 * Feed both to aat_segment_mean_pool (n_rows = sum of T_b, or AAT_POOL_ROWS_FROM_DEVICE). */
int aat_utterance_frame_csr(aat_ctx *ctx, const aat_plan *plan, const int64_t *seg_start_dev,
                            const int32_t *seg_count_dev, const int64_t *utt_seg_off_dev, int64_t *seg_off_dev,
                            int64_t *n_seg_dev, void *stream);

/* ------------------------------------------------------------------ K4: ragged mean-pool
 * Replaces `torch.cat([x.mean(dim=1, keepdim=True).to(float32) for x in embs], dim=1)`
 * (ref:scripts/mean_hubert_embeddings.py:19-20) on the packed layout
 *   emb_dev [n_rows, dim] row-major (AAT_F32 / AAT_F16 / AAT_BF16) + seg_off_dev [S+1] (frame units).
 * plan      : optional.  The kernel needs scratch for the partial sums of segments that straddle two CTAs; a plan
 *             owns such a block (like every other device-side state of a step), so launches that name different
 *             plans may run concurrently on different streams.  NULL = the context's block: such launches are
 *             ordered against each other with an event (whatever their streams), and must not be captured into a
 *             CUDA graph concurrently with other plan-less launches.
 * out_dev   : float32 [S, dim]; an empty segment yields NaN, as torch's mean does
 * n_rows    : rows of emb_dev; rows outside [seg_off[0], seg_off[S]) belong to no segment and are ignored
 * n_seg     : S when n_seg_dev is NULL; otherwise an upper bound and S is read from n_seg_dev[0]
 * colsum_dev: optional float64 [dim+1]: column sums over the S pooled vectors and, last, S itself
 *             (input of the dataset-level mean allreduce)
 * flags     : bit set of aat_pool_flags
 * One pass over emb; dim * sizeof(element) must be a multiple of 16 and emb_dev 16-byte aligned. */
typedef enum aat_pool_flags {
    AAT_POOL_ACCUMULATE = 1,       /* this batch's column sums are ADDED to colsum_dev (running totals over batches
                                      without a separate kernel); default: colsum_dev is overwritten */
    AAT_POOL_EMB_READY = 2,        /* the caller vouches that the kernel enqueued immediately before this call on
                                      `stream` does not write emb_dev (true when it is aat_boundaries): the kernel may
                                      then request embedding rows before it waits for that predecessor.  Without the
                                      flag nothing is read before the wait, so any producer of emb_dev may precede
                                      the call, including kernels that trigger their dependents early */
    AAT_POOL_ROWS_FROM_DEVICE = 4, /* n_rows is an upper bound (the allocation); the rows the CSR covers are read
                                      from n_seg_dev[1] as written by aat_boundaries / aat_segment_frame_csr, and
                                      nothing beyond them is streamed */
    AAT_POOL_SHARE_SMS = 8         /* several batches are in flight (other plans on other streams): launch ONE CTA per
                                      SM instead of two, so that the pool kernels of two batches, or a pool kernel and a
                                      log-mel CTA, fit on an SM together.  The kernel alone is ~8 % slower that way,
                                      a pipelined schedule 2-3 % faster (profiles/r2_pipeline_ab.txt).  The order in
                                      which a segment's rows are added follows the CTA tiles, and the flag halves the
                                      number of tiles: a mean may differ in the last float32 bit from a launch without
                                      the flag (either way deterministic, and far inside the 1e-5 bound) */
} aat_pool_flags;
int aat_segment_mean_pool(aat_ctx *ctx, const aat_plan *plan, const void *emb_dev, int emb_dtype, int64_t n_rows,
                          int32_t dim, const int64_t *seg_off_dev, int64_t n_seg, const int64_t *n_seg_dev,
                          float *out_dev, double *colsum_dev, int flags, void *stream);

/* ------------------------------------------------------------------ one step in one call
 * What the reference's offline job does per item — tokenize (ref:scripts/audio_tokenization.py:33-41, with the z-score of
 * ref:scripts/audio_tokenization_melspec.py:40 when `znorm`) and pool its embeddings
 * (ref:scripts/mean_hubert_embeddings.py:16-23) — for a whole batch:
 * aat_logmel (with the fused z-score when `znorm` is non-zero: statistics by aat_normalize into bufs->znorm_stats
 * first) -> aat_boundaries (with the packed frame CSR) -> aat_segment_mean_pool, enqueued back to back on `stream`:
 * exactly the launches the separate calls make, for loops that drive many steps per second from an interpreted host
 * language (three or four foreign calls per step cost more host time than one; aat_b200/pipeline.py).
 * bufs: the plan-sized device buffers of the separate entry points, same meaning and sizes; minima / minima_count /
 * utt_seg_off / znorm_stats may be NULL (znorm_stats only when znorm is 0). */
typedef struct aat_step_buffers {
    float *mel;            /* aat_logmel: mel_dev                       */
    float *amp;            /* aat_logmel: amp_dev (required here)       */
    int64_t *seg_start;    /* aat_boundaries ...                        */
    int64_t *seg_len;
    int32_t *seg_count;
    int64_t *minima;
    int32_t *minima_count;
    int32_t *status;
    int64_t *seg_off;
    int64_t *n_seg;        /* int64 [2]                                 */
    int64_t *utt_seg_off;
    double *znorm_stats;   /* float64 [2 * n_utts]                      */
} aat_step_buffers;
int aat_tokenize_and_pool(aat_ctx *ctx, const aat_plan *plan, const aat_step_buffers *bufs, const void *wave_dev,
                          int wave_dtype, int znorm, const void *emb_dev, int emb_dtype, int64_t n_rows, int32_t dim,
                          float *out_dev, int64_t out_capacity, double *colsum_dev, int pool_flags, void *stream);

/* acc_dev[0..dim] += colsum_dev[0..dim] (float64), for accumulating over batches before the allreduce. */
int aat_colsum_accumulate(aat_ctx *ctx, double *acc_dev, const double *colsum_dev, int32_t dim, void *stream);
/* mean_dev[d] = float32(acc_dev[d] / acc_dev[dim]) — after the SUM allreduce of acc_dev over ranks. */
int aat_colsum_finalize(aat_ctx *ctx, const double *acc_dev, int32_t dim, float *mean_dev, void *stream);

/* ------------------------------------------------------------------ callers either side of the path
 * (SURVEY.md section 8f rows N1, N2).  Device entry points, asynchronous on `stream`. */

typedef enum aat_norm_mode {
    AAT_NORM_ZSCORE = 0, /* (x - mean) / (std + 1e-6), float64: ref:scripts/audio_tokenization_melspec.py:40,
                            ref:src/aat/training/collate.py:135,138,152 */
    AAT_NORM_W2V2 = 1    /* (x - mean) / sqrt(var + 1e-7), float32: ref:src/aat/training/collate.py:301 ->
                            Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm */
} aat_norm_mode;

/* Per-utterance waveform normalisation on the packed layout.  Statistics are accumulated in float64.
 * out_dev   : packed like the input, AAT_F32 or AAT_F64; NULL = statistics only
 * stats_dev : optional float64 [2 * n_utts]: (mean, population variance) per utterance */
int aat_normalize(aat_ctx *ctx, const aat_plan *plan, const void *wave_dev, int in_dtype, int mode, void *out_dev,
                  int out_dtype, double *stats_dev, void *stream);

/* The same normalisation written straight into the feature extractor's padded layout
 * (ref:src/aat/training/collate.py:301-304, the processor call with padding=True):
 * out_dev [n_utts, n_max] float32 = normalised samples, zeros behind each utterance's end;
 * mask_dev (optional) [n_utts, n_max] int32 = the processor's attention_mask (the feature extractor returns int32).  Every utterance must fit n_max. */
int aat_normalize_padded(aat_ctx *ctx, const aat_plan *plan, const void *wave_dev, int in_dtype, int mode, float *out_dev,
                         int64_t n_max, int32_t *mask_dev, double *stats_dev, void *stream);

/* `_make_padded_segments_boarders` (ref:src/aat/training/collate.py:242-253) on the output of aat_boundaries:
 * boarders_dev [n_utts, s_max] = cumulative segment ends (the collator's `frames_boarders`, :158), zero padded;
 * mask_dev [n_utts, s_max] = 1 for real segments.  status_dev [n_utts] gets AAT_ERR_CAPACITY when an utterance
 * has more than s_max segments. */
int aat_pad_segment_boarders(aat_ctx *ctx, const aat_plan *plan, const int64_t *seg_len_dev,
                             const int32_t *seg_count_dev, int64_t s_max, int64_t *boarders_dev, int64_t *mask_dev,
                             int32_t *status_dev, void *stream);

/* The collator's waveform scatter (ref:src/aat/training/collate.py:321-335):
 * out_dev [n_utts, s_max, max_frames] = wave_padded_dev[b, prev:boarder] left aligned, zero elsewhere;
 * mask_dev (optional) likewise with ones.  wave_padded_dev is [n_utts, n_max] float32 (the feature extractor's
 * padded `input_values`).  status_dev [n_utts] gets AAT_ERR_INVALID where the reference would raise
 * (non-increasing boarders, a segment longer than max_frames or running past n_max). */
int aat_scatter_segments(aat_ctx *ctx, const float *wave_padded_dev, int64_t n_max, int32_t n_utts,
                         const int64_t *boarders_dev, int64_t s_max, int64_t max_frames, float *out_dev,
                         float *mask_dev, int32_t *status_dev, void *stream);

/* The collator's log-mel scatter (ref:src/aat/training/collate.py:337-342):
 * out_dev [n_utts, s_max, n_mels, max_items] = mel_b[:, prev // hop : boarder // hop], zero elsewhere;
 * mel_dev is the packed log-mel of the plan. */
int aat_scatter_mel_segments(aat_ctx *ctx, const aat_plan *plan, const float *mel_dev, const int64_t *boarders_dev,
                             int64_t s_max, int64_t max_items, float *out_dev, int32_t *status_dev, void *stream);

/* The same scatter from mel blocks the caller describes instead of a plan's packed layout: utterance b's mel has
 * n_mels rows of mel_frames_dev[b] columns, row r starting at element mel_elem_off_dev[b] + r * mel_row_stride_dev[b]
 * of mel_dev — e.g. the cropped mels of the collator's n-word path (ref:src/aat/training/collate.py:208-212), which
 * are column slices (views) of the full ones.  All three arrays are int64 [n_utts]. */
int aat_scatter_mel_tiles(aat_ctx *ctx, int32_t n_utts, const float *mel_dev, const int64_t *mel_elem_off_dev,
                          const int64_t *mel_frames_dev, const int64_t *mel_row_stride_dev, const int64_t *boarders_dev,
                          int64_t s_max, int64_t max_items, float *out_dev, int32_t *status_dev, void *stream);

/* Masked mean over the valid frames of the padded layout: the
 * `SegmentProjectionEnum.mean` branch that the reference leaves as NotImplementedError
 * (ref:src/aslm/modeling_aslm.py:258-259), under the frame mask built by encode_audio (:195-218).
 * emb_dev [n_rows, seq_len, dim] (F32/F16/BF16), mask_dev [n_rows, seq_len] int64 (non-zero = valid)
 * -> out_dev [n_rows, dim] float32 = sum of the valid frames / their count (zeros when none is valid);
 * row_mask_dev (optional) [n_rows] int64 = 1 where the row has a valid frame (the segment-level mask). */
int aat_masked_mean_pool(aat_ctx *ctx, const void *emb_dev, int emb_dtype, int64_t n_rows, int64_t seq_len, int32_t dim,
                         const int64_t *mask_dev, float *out_dev, int64_t *row_mask_dev, void *stream);

/* ------------------------------------------------------------------ synthetic inputs, generated on the device
 * For the benchmark's dataset-scale job (BASELINE config 5: 1000 audio-hours of DISTINCT utterances) and for tests:
 * inputs of the path, never results.  Counter-based (Philox 4x32-10): sample i of utterance u is a pure function
 * of (seed_base + utt_index_base + u, i), so any rank can generate any shard.  Recipe of SURVEY.md section 8d, the
 * one aat_b200/synth.py implements on the host: Gaussian noise times an envelope of Hann-shaped bursts of
 * U(80,600) ms, amplitude U(0.3,1.0), separated by pauses of U(30,250) ms at a 1e-3 floor.
 * wave_dev      : packed float32 waveform of the plan (aat_plan_total_samples entries)
 * workspace_dev : aat_synth_workspace_bytes(plan) bytes, 16-byte aligned (burst tables) */
int64_t aat_synth_workspace_bytes(const aat_plan *plan);
int aat_synth_waveforms(aat_ctx *ctx, const aat_plan *plan, uint64_t seed_base, int64_t utt_index_base, float *wave_dev,
                        void *workspace_dev, void *stream);
/* out_dev[i] ~ N(0, 1), i < n: random-init HuBERT-shaped embeddings.  out_dev 16-byte aligned. */
int aat_synth_normal(aat_ctx *ctx, float *out_dev, int64_t n, uint64_t seed, void *stream);

/* ------------------------------------------------------------------ host-buffer entry points
 * Same operations for callers that hold numpy arrays, exactly like the reference's methods:
 * the library stages host<->device copies in its own scratch and synchronises. */

/* get_melspec (ref:src/aat/tokenizer.py:107): wave_host [n_samples] -> mel_host (n_mels, 1+n/hop) float32 */
int aat_host_logmel(aat_ctx *ctx, const void *wave_host, int wave_dtype, int64_t n_samples, float *mel_host);

/* find_amplitude_minimas (ref:src/aat/tokenizer.py:55): mel_host (n_mels, n_frames) C-contiguous float32
 * -> minima_host (capacity n_frames), *n_minima_host */
int aat_host_find_minimas(aat_ctx *ctx, const float *mel_host, int64_t n_frames, int64_t *minima_host,
                          int64_t *n_minima_host);

/* process_segments_boarders (ref:src/aat/tokenizer.py:141) */
int aat_host_process_boarders(aat_ctx *ctx, int64_t n_samples, const int64_t *boarders_host, int64_t n_boarders,
                              int64_t *seg_start_host, int64_t *seg_len_host, int64_t capacity,
                              int64_t *n_segments_host, int32_t *padded_tail_host);

/* tokenize / pretokenize+process_segments_boarders (ref:src/aat/tokenizer.py:121-200) for one utterance.
 * mel_in_host  : optional precomputed mel (the `melspec=` argument); NULL = compute from wave
 * mel_out_host : optional, receives the mel that was used
 * minima_host  : optional (capacity n_frames)
 * padded_tail_host : optional; 1 when the last segment is the zero-padded tail */
int aat_host_tokenize(aat_ctx *ctx, const void *wave_host, int wave_dtype, int64_t n_samples,
                      const float *mel_in_host, float *mel_out_host, int64_t *minima_host,
                      int64_t *n_minima_host, int64_t *seg_start_host, int64_t *seg_len_host, int64_t capacity,
                      int64_t *n_segments_host, int32_t *padded_tail_host);

/* mean_hubert_embeddings pooling (ref:scripts/mean_hubert_embeddings.py:19-20) on host buffers */
int aat_host_mean_pool(aat_ctx *ctx, const void *emb_host, int emb_dtype, int64_t n_rows, int32_t dim,
                       const int64_t *seg_off_host, int64_t n_seg, float *out_host, double *colsum_host);

/* The same on the reference's own argument: a LIST of per-segment tensors [1, n_i, dim] in host memory
 * (ref:scripts/mean_hubert_embeddings.py:18: what torch.load returns for one file).  seg_ptrs_host[i] points at the
 * n_i x dim contiguous elements of segment i, seg_rows_host[i] = n_i.  Every tensor is copied once, straight into the
 * library's pinned staging buffer (no concatenation on the host first). */
int aat_host_mean_pool_list(aat_ctx *ctx, const void *const *seg_ptrs_host, const int64_t *seg_rows_host, int64_t n_seg,
                            int emb_dtype, int32_t dim, float *out_host, double *colsum_host);

/* Upper bound on the segments one utterance can produce (slot capacity used by plans). */
int64_t aat_segment_capacity(const aat_config *cfg, int64_t n_samples);
/* 1 + n_samples / hop (TF:audio_utils.py:778 with centre padding). */
int64_t aat_num_mel_frames(const aat_config *cfg, int64_t n_samples);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* AAT_B200_H */
