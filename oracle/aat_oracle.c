/*
 * ORACLE (test infrastructure, not product code) — plain C restatement of the
 * order-sensitive parts of the reference's tokenization front end.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library.  Build: `make -C oracle` -> oracle/libaat_oracle.so
 * (gcc -O2 -ffp-contract=off: no FMA contraction, strict IEEE float32).
 *
 * Reference sites followed
 *   ref:src/aat/tokenizer.py:67      amp   = -10 * melspec.mean(axis=0)   (numpy: sequential fp32 down the rows, /M)
 *   ref:src/aat/tokenizer.py:71-75   cs    = cumsum(amp) (sequential fp32); rm[i] = (cs[i+N]-cs[i]) / float(N)
 *   ref:src/aat/tokenizer.py:82-85   argrelextrema(rm, a > b + 1e-5), order 1, mode='clip'
 *                                    (SP:signal/_peak_finding.py:66-79: end points compare against themselves)
 *   ref:src/aat/tokenizer.py:90      keep rm[m] > max_amplitude_for_minima
 *   ref:src/aat/tokenizer.py:141-183 merge-small / split-big state machine, np.split clamp semantics, padded tail
 *   ref:scripts/mean_hubert_embeddings.py:19-20  per-segment mean over frames
 *   TF:audio_utils.py:769-830        log-mel (here with a naive O(n^2) long-double DFT: an FFT-independent check)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* amp[t] = -10 * ((sum_r mel[r*ld + t]) / M), all float32, rows added in order r = 0..M-1 */
ORC_API void orc_amp_curve(const float *mel, int64_t n_mels, int64_t T, int64_t ld, float *amp)
{
    for (int64_t t = 0; t < T; ++t) {
        volatile float acc = mel[t];
        for (int64_t r = 1; r < n_mels; ++r)
            acc = acc + mel[r * ld + t];
        volatile float mean = acc / (float)n_mels;
        amp[t] = -10.0f * mean;
    }
}

ORC_API void orc_cumsum_f32(const float *x, int64_t n, float *cs)
{
    volatile float s = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        s = (i == 0) ? x[0] : s + x[i];
        cs[i] = s;
    }
}

/* Returns the number of minima written (indices into rm, i.e. mel-frame indices). */
ORC_API int64_t orc_find_minimas(const float *mel, int64_t n_mels, int64_t T, int64_t ld,
                                 int64_t running_mean_points, float max_amplitude,
                                 int64_t *minima, float *amp_out, float *cs_out, float *rm_out)
{
    float *amp = (float *)malloc(sizeof(float) * (size_t)(T > 0 ? T : 1));
    float *cs = (float *)malloc(sizeof(float) * (size_t)(T > 0 ? T : 1));
    orc_amp_curve(mel, n_mels, T, ld, amp);
    orc_cumsum_f32(amp, T, cs);
    int64_t n = running_mean_points;
    int64_t L = T - n;
    if (L < 0) L = 0;
    float *rm = (float *)malloc(sizeof(float) * (size_t)(L > 0 ? L : 1));
    for (int64_t i = 0; i < L; ++i) {
        volatile float d = cs[i + n] - cs[i];
        rm[i] = d / (float)n;
    }
    const float eps = 1e-5f;
    int64_t count = 0;
    for (int64_t i = 1; i + 1 < L; ++i) {
        volatile float right = rm[i + 1] + eps;
        volatile float left = rm[i - 1] + eps;
        if (rm[i] > right && rm[i] > left && rm[i] > max_amplitude)
            minima[count++] = i;
    }
    if (amp_out) memcpy(amp_out, amp, sizeof(float) * (size_t)T);
    if (cs_out) memcpy(cs_out, cs, sizeof(float) * (size_t)T);
    if (rm_out) memcpy(rm_out, rm, sizeof(float) * (size_t)L);
    free(amp); free(cs); free(rm);
    return count;
}

/* Returns the number of segments, or -1 when capacity is too small, or -2 when the
 * reference would raise (tail longer than min_frames).  padded_tail: last entry is the zero-padded tail. */
ORC_API int64_t orc_state_machine(int64_t n_samples, const int64_t *boarders, int64_t n_boarders,
                                  int64_t min_frames, int64_t max_frames,
                                  int64_t *starts, int64_t *lengths, int64_t capacity, int32_t *padded_tail)
{
    int64_t count = 0, prev = 0;
#define EMIT(s, l) do { if (count >= capacity) return -1; starts[count] = (s); lengths[count] = (l); ++count; } while (0)
    for (int64_t bi = 0; bi < n_boarders; ++bi) {
        int64_t b = boarders[bi];
        int64_t len = b - prev;
        if (len < min_frames) continue;
        if (len > max_frames) {
            int64_t k = len / max_frames;
            int64_t gap = len - k * max_frames;
            int64_t n_cuts = k;
            int64_t last_cut = k * max_frames;
            if (gap == 0) n_cuts = k - 1;
            else if (gap < min_frames) last_cut = len - min_frames;
            int64_t lo = 0;
            for (int64_t j = 0; j < n_cuts; ++j) {
                int64_t c = (j == k - 1) ? last_cut : (j + 1) * max_frames;
                int64_t a0 = lo < len ? lo : len, a1 = c < len ? c : len;
                EMIT(prev + a0, a1 > a0 ? a1 - a0 : 0);
                lo = c;
            }
            int64_t a0 = lo < len ? lo : len;
            EMIT(prev + a0, len - a0);
        } else {
            EMIT(prev, len);
        }
        prev = b;
    }
    *padded_tail = (prev != n_samples);
    if (*padded_tail) {
        if (n_samples - prev > min_frames) return -2;
        EMIT(prev, min_frames);
    }
#undef EMIT
    return count;
}

/* pooled[s][d] = mean over rows [off[s], off[s+1]) — float32 sequential accumulate, true division */
ORC_API void orc_mean_pool_f32(const float *emb, int64_t D, const int64_t *off, int64_t S, float *out)
{
    for (int64_t s = 0; s < S; ++s) {
        int64_t a = off[s], b = off[s + 1];
        for (int64_t d = 0; d < D; ++d) {
            volatile float acc = 0.0f;
            for (int64_t r = a; r < b; ++r) acc = acc + emb[r * D + d];
            out[s * D + d] = acc / (float)(b - a);
        }
    }
}

ORC_API void orc_mean_pool_f64(const float *emb, int64_t D, const int64_t *off, int64_t S, double *out)
{
    for (int64_t s = 0; s < S; ++s) {
        int64_t a = off[s], b = off[s + 1];
        for (int64_t d = 0; d < D; ++d) {
            double acc = 0.0;
            for (int64_t r = a; r < b; ++r) acc += (double)emb[r * D + d];
            out[s * D + d] = acc / (double)(b - a);
        }
    }
}

static int64_t reflect_idx(int64_t p, int64_t n)
{
    if (n == 1) return 0;
    int64_t period = 2 * (n - 1);
    int64_t m = p % period;
    if (m < 0) m += period;
    return m < n ? m : period - m;
}

/* Log-mel with a naive DFT evaluated in long double — checks the product's FFT against
 * the definition rather than against another FFT.  wave: float64; mel_filters: (n_bins, n_mels)
 * row-major float64; out: (n_mels, T) float32 with T = 1 + n/hop. */
ORC_API void orc_logmel_naive(const double *wave, int64_t n, const double *window, int64_t n_fft,
                              int64_t hop, const double *mel_filters, int64_t n_mels, float *out)
{
    int64_t pad = n_fft / 2, n_bins = n_fft / 2 + 1;
    int64_t T = 1 + n / hop;
    long double *cs = (long double *)malloc(sizeof(long double) * (size_t)n_fft);
    long double *sn = (long double *)malloc(sizeof(long double) * (size_t)n_fft);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int64_t j = 0; j < n_fft; ++j) {
        cs[j] = cosl(two_pi * (long double)j / (long double)n_fft);
        sn[j] = sinl(two_pi * (long double)j / (long double)n_fft);
    }
    double *frame = (double *)malloc(sizeof(double) * (size_t)n_fft);
    double *power = (double *)malloc(sizeof(double) * (size_t)n_bins);
    for (int64_t t = 0; t < T; ++t) {
        for (int64_t j = 0; j < n_fft; ++j) {
            double v = wave[reflect_idx(t * hop + j - pad, n)];
            frame[j] = v * window[j];
        }
        for (int64_t k = 0; k < n_bins; ++k) {
            long double re = 0.0L, im = 0.0L;
            for (int64_t j = 0; j < n_fft; ++j) {
                int64_t ph = (k * j) % n_fft;
                re += (long double)frame[j] * cs[ph];
                im -= (long double)frame[j] * sn[ph];
            }
            float ref = (float)(double)re, imf = (float)(double)im; /* complex64 store, TF:audio_utils.py:781,803 */
            double mag = hypot((double)ref, (double)imf);            /* np.abs(.., dtype=float64) */
            power[k] = mag * mag;                                    /* ** 2.0 */
        }
        for (int64_t m = 0; m < n_mels; ++m) {
            double acc = 0.0;
            for (int64_t k = 0; k < n_bins; ++k) acc += mel_filters[k * n_mels + m] * power[k];
            if (acc < 1e-10) acc = 1e-10;
            out[m * T + t] = (float)log10(acc);
        }
    }
    free(cs); free(sn); free(frame); free(power);
}
