// K1+K2: batched framed STFT -> power spectrum -> sparse mel projection -> log10.
//
// Replaces get_melspec -> transformers.audio_utils.spectrogram (ref:src/aat/tokenizer.py:107-119,
// TF:audio_utils.py:769-830).  The reference computes in float64, rounds the spectrum to complex64,
// takes |.|^2 in float64, projects with a float64 dgemm and stores float32(log10).  This kernel keeps
// exactly that precision ladder so the float32 log-mel agrees with the reference to the last bit in
// all but ~1e-8 of the elements (the float64 DFT round-off is 9 orders of magnitude below float32
// resolution); B200 runs FP64 at half the FP32 rate, so the cost is ~2x ALU, not a different answer.
//
// Decomposition (n_fft = 400 = 20 x 20, not a power of two):
//   * two real frames are packed into one complex 400-point transform  z = frame_a + i * frame_b
//   * 400-point DFT = Cooley-Tukey 20 x 20: 20 threads per frame pair, each thread runs a 20-point
//     DFT entirely in registers (prime-factor 4 x 5, so no inner twiddles), one shared-memory
//     transpose with the W_400 twiddles between the two passes
//   * split: X_a[k] = (Z[k] + conj Z[400-k]) / 2,  X_b[k] = (Z[k] - conj Z[400-k]) / 2i
//     (the 1/2 is folded into the window table: an exact power-of-two scaling)
//   * power (float64 of the float32-rounded re/im), sparse triangular mel (<= 2 filters per bin,
//     388 non-zeros: a gather, not a dense contraction -> CUDA cores, no tensor cores), log10, float32
//   * optional fused epilogue: amp[t] = -10 * mean_m(mel[m][t]) in numpy's sequential float32 order
//     (ref:src/aat/tokenizer.py:67), so the boundary kernel does not have to re-read the mel.
//
// One CTA = 16 consecutive frames of one utterance (8 frame pairs x 20 threads = 160 threads).
#include "aat_internal.cuh"

namespace aat {

namespace {

constexpr int kFrames = kMelFramesPerTile;   // 16
constexpr int kPairs = kFrames / 2;          // 8
constexpr int kThreads = kPairs * 20;        // 160
constexpr int kRow = 21;                     // padded row (double2 units): 21 is odd -> conflict-free columns
constexpr int kPairStride = 20 * kRow;       // 420 double2; 420 % 8 == 4 keeps neighbouring pairs on distinct banks
constexpr int kPowStride = kBins;            // 201 doubles (odd)

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 20-point forward DFT in registers, Good-Thomas 4 x 5:
//   n = (5 n1 + 4 n2) mod 20,  k = (5 k1 + 16 k2) mod 20,  X[k] = sum x[n] W4^(n1 k1) W5^(n2 k2)
__device__ __forceinline__ void dft20(double2 (&v)[20])
{
    // five radix-4 butterflies over n1 (stride 5), in place: slot (5 k1 + 4 n2) % 20
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        const int i0 = (4 * n2) % 20, i1 = (5 + 4 * n2) % 20, i2 = (10 + 4 * n2) % 20, i3 = (15 + 4 * n2) % 20;
        double2 s0 = cadd(v[i0], v[i2]), d0 = csub(v[i0], v[i2]);
        double2 s1 = cadd(v[i1], v[i3]), d1 = csub(v[i1], v[i3]);
        v[i0] = cadd(s0, s1);
        v[i2] = csub(s0, s1);
        v[i1] = make_double2(d0.x + d1.y, d0.y - d1.x); // d0 - i d1
        v[i3] = make_double2(d0.x - d1.y, d0.y + d1.x); // d0 + i d1
    }
    // four radix-5 butterflies over n2 (stride 4), result k2 lands in slot (5 k1 + 16 k2) % 20
    constexpr double c1 = 0.30901699437494742410229341718282;  // cos(2 pi / 5)
    constexpr double c2 = -0.80901699437494742410229341718282; // cos(4 pi / 5)
    constexpr double s1 = 0.95105651629515357211643933337938;  // sin(2 pi / 5)
    constexpr double s2 = 0.58778525229247312916870595463907;  // sin(4 pi / 5)
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        const int b = 5 * k1;
        const int j0 = b % 20, j1 = (b + 4) % 20, j2 = (b + 8) % 20, j3 = (b + 12) % 20, j4 = (b + 16) % 20;
        double2 x0 = v[j0];
        double2 t1 = cadd(v[j1], v[j4]), t3 = csub(v[j1], v[j4]);
        double2 t2 = cadd(v[j2], v[j3]), t4 = csub(v[j2], v[j3]);
        double2 m1 = make_double2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
        double2 m2 = make_double2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
        double2 n1 = make_double2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
        double2 n2 = make_double2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
        // k2 = 0..4 -> slots (b + 16 k2) % 20 = j0, j4, j3, j2, j1
        v[j0] = make_double2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
        v[j4] = make_double2(m1.x + n1.y, m1.y - n1.x); // k2 = 1: m1 - i n1
        v[j3] = make_double2(m2.x + n2.y, m2.y - n2.x); // k2 = 2: m2 - i n2
        v[j2] = make_double2(m2.x - n2.y, m2.y + n2.x); // k2 = 3: m2 + i n2
        v[j1] = make_double2(m1.x - n1.y, m1.y + n1.x); // k2 = 4: m1 + i n1
    }
}

__device__ __forceinline__ int64_t reflect_index(int64_t g, int64_t n)
{
    if (g >= 0 && g < n) return g;
    if (n == 1) return 0;
    const int64_t period = 2 * (n - 1);
    int64_t m = g % period;
    if (m < 0) m += period;
    return m < n ? m : period - m;
}

struct LogmelParams {
    const void *wave;
    float *mel;
    float *amp;
    const int64_t *n_samples;
    const int64_t *wave_off;
    const int64_t *frame_off;
    const int32_t *tile_utt;
    const int32_t *tile_first;
    const double *window_half;
    const double2 *twiddle;
    const int *mel_row_start;
    const int *mel_bin;
    const double *mel_weight;
    int hop;
    int n_mels;
    int nnz;
    int stage_len; // (kFrames - 1) * hop + 400
};

template <typename WaveT>
__global__ void __launch_bounds__(kThreads) logmel_kernel(const LogmelParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // region A: wave staging + window, later overlaid by the power spectra
    double *s_wave = reinterpret_cast<double *>(smem_raw);
    double *s_win = s_wave + p.stage_len;
    double *s_pow = reinterpret_cast<double *>(smem_raw);
    size_t region_a = sizeof(double) * (size_t)max(p.stage_len + kNfft, kFrames * kPowStride);
    region_a = (region_a + 15) & ~size_t(15);
    double2 *s_ex = reinterpret_cast<double2 *>(smem_raw + region_a);
    double2 *s_tw = s_ex + kPairs * kPairStride;
    double *s_mw = reinterpret_cast<double *>(s_tw + 400);
    int *s_mbin = reinterpret_cast<int *>(s_mw + p.nnz);
    int *s_mrow = s_mbin + p.nnz;
    float *s_mel = reinterpret_cast<float *>(s_mrow + p.n_mels + 1);

    const int tid = threadIdx.x;
    const int utt = p.tile_utt[blockIdx.x];
    const int tile = blockIdx.x - p.tile_first[utt];
    const int64_t n = p.n_samples[utt];
    const int64_t T = 1 + n / p.hop;
    const int64_t f0 = (int64_t)tile * kFrames;
    const WaveT *wave = reinterpret_cast<const WaveT *>(p.wave) + p.wave_off[utt];

    // ---- stage 0: constants and the tile's samples (reflect padding, TF:audio_utils.py:769-771) ----
    for (int i = tid; i < kNfft; i += kThreads) {
        s_win[i] = p.window_half[i];
        s_tw[i] = p.twiddle[i];
    }
    for (int i = tid; i < p.nnz; i += kThreads) {
        s_mw[i] = p.mel_weight[i];
        s_mbin[i] = p.mel_bin[i];
    }
    for (int i = tid; i <= p.n_mels; i += kThreads) s_mrow[i] = p.mel_row_start[i];
    {
        const int64_t g0 = f0 * p.hop - kNfft / 2;
        for (int i = tid; i < p.stage_len; i += kThreads)
            s_wave[i] = (double)wave[reflect_index(g0 + i, n)];
    }
    __syncthreads();

    const int pair = tid / 20;
    const int lane20 = tid - pair * 20;
    double2 *ex = s_ex + pair * kPairStride;

    // ---- stage 1: thread n2 transforms x[20 n1 + n2] over n1, applies W_400^(n2 k1) ----
    {
        double2 v[20];
        const double *wa = s_wave + (2 * pair) * p.hop + lane20;
        const double *wb = wa + p.hop;
#pragma unroll
        for (int n1 = 0; n1 < 20; ++n1) {
            const double w = s_win[20 * n1 + lane20];
            v[n1] = make_double2(wa[20 * n1] * w, wb[20 * n1] * w);
        }
        dft20(v);
        ex[lane20] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < 20; ++k1) ex[k1 * kRow + lane20] = cmul(v[k1], s_tw[k1 * 20 + lane20]);
    }
    __syncthreads();

    // ---- stage 2: thread k1 transforms row k1 over n2; Z[k1 + 20 k2] goes back into its own row ----
    {
        double2 v[20];
        double2 *row = ex + lane20 * kRow;
#pragma unroll
        for (int n2 = 0; n2 < 20; ++n2) v[n2] = row[n2];
        dft20(v);
#pragma unroll
        for (int k2 = 0; k2 < 20; ++k2) row[k2] = v[k2];
    }
    __syncthreads();

    // ---- split into the two real spectra, round to complex64, power in float64 ----
    for (int item = tid; item < kPairs * kBins; item += kThreads) {
        const int pr = item / kBins;
        const int k = item - pr * kBins;
        const double2 *e = s_ex + pr * kPairStride;
        const int kk = (k == 0) ? 0 : kNfft - k;
        const double2 z = e[(k % 20) * kRow + k / 20];
        const double2 y = e[(kk % 20) * kRow + kk / 20];
        // X_a = (z + conj y), X_b = (z - conj y) / i   (the 1/2 lives in the window table)
        const float ar = (float)(z.x + y.x), ai = (float)(z.y - y.y);
        const float br = (float)(z.y + y.y), bi = (float)(y.x - z.x);
        s_pow[(2 * pr) * kPowStride + k] = (double)ar * (double)ar + (double)ai * (double)ai;
        s_pow[(2 * pr + 1) * kPowStride + k] = (double)br * (double)br + (double)bi * (double)bi;
    }
    __syncthreads();

    // ---- mel projection (sparse rows), floor, log10, float32 store ----
    float *mel_out = p.mel + (size_t)p.n_mels * p.frame_off[utt];
    for (int item = tid; item < p.n_mels * kFrames; item += kThreads) {
        const int m = item / kFrames;
        const int f = item - m * kFrames;
        const double *pw = s_pow + f * kPowStride;
        double acc = 0.0;
        for (int j = s_mrow[m]; j < s_mrow[m + 1]; ++j) acc = fma(s_mw[j], pw[s_mbin[j]], acc);
        const float out = (float)log10(fmax(acc, 1e-10));
        if (f0 + f < T) mel_out[(size_t)m * T + f0 + f] = out;
        s_mel[f * (p.n_mels + 1) + m] = out;
    }

    // ---- fused amplitude curve: numpy's mean(axis=0) adds the rows in order in float32 ----
    if (p.amp != nullptr) {
        __syncthreads();
        if (tid < kFrames && f0 + tid < T) {
            const float *col = s_mel + tid * (p.n_mels + 1);
            float acc = col[0];
            for (int m = 1; m < p.n_mels; ++m) acc = __fadd_rn(acc, col[m]);
            const float mean = __fdiv_rn(acc, (float)p.n_mels);
            p.amp[p.frame_off[utt] + f0 + tid] = __fmul_rn(-10.0f, mean);
        }
    }
}

size_t logmel_smem_bytes(int hop, int n_mels, int nnz)
{
    const int stage_len = (kFrames - 1) * hop + kNfft;
    size_t region_a = sizeof(double) * (size_t)((stage_len + kNfft) > kFrames * kPowStride ? (stage_len + kNfft)
                                                                                              : kFrames * kPowStride);
    region_a = (region_a + 15) & ~size_t(15);
    size_t bytes = region_a;
    bytes += sizeof(double2) * (kPairs * kPairStride + 400);
    bytes += sizeof(double) * nnz + sizeof(int) * (nnz + n_mels + 1);
    bytes += sizeof(float) * kFrames * (n_mels + 1);
    return bytes;
}

} // namespace

int launch_logmel(aat_ctx *ctx, const aat_plan *plan, const void *wave, int wave_dtype, float *mel, float *amp,
                  cudaStream_t stream)
{
    AAT_REQUIRE(wave_dtype == AAT_F32 || wave_dtype == AAT_F64, AAT_ERR_UNSUPPORTED,
                "aat_logmel: waveform dtype must be AAT_F32 or AAT_F64 (got %d)", wave_dtype);
    if (plan->mel_tiles == 0) return AAT_OK;
    LogmelParams p{};
    p.wave = wave;
    p.mel = mel;
    p.amp = amp;
    p.n_samples = plan->d_n_samples;
    p.wave_off = plan->d_wave_off;
    p.frame_off = plan->d_frame_off;
    p.tile_utt = plan->d_tile_utt;
    p.tile_first = plan->d_tile_first;
    p.window_half = ctx->window_half;
    p.twiddle = ctx->twiddle;
    p.mel_row_start = ctx->mel.row_start;
    p.mel_bin = ctx->mel.bin;
    p.mel_weight = ctx->mel.weight;
    p.hop = ctx->cfg.hop_length;
    p.n_mels = ctx->mel.n_mels;
    p.nnz = ctx->mel.nnz;
    p.stage_len = (kFrames - 1) * p.hop + kNfft;
    const size_t smem = logmel_smem_bytes(p.hop, p.n_mels, p.nnz);
    ProfileScope prof(ctx, AAT_K_LOGMEL, stream);
    if (wave_dtype == AAT_F32) {
        AAT_CUDA_CHECK(cudaFuncSetAttribute(logmel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        logmel_kernel<float><<<plan->mel_tiles, kThreads, smem, stream>>>(p);
    } else {
        AAT_CUDA_CHECK(cudaFuncSetAttribute(logmel_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        logmel_kernel<double><<<plan->mel_tiles, kThreads, smem, stream>>>(p);
    }
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat
