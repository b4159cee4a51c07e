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
//   * power (float64 of the float32-rounded re/im), banded triangular mel (each filter touches 2-18
//     consecutive bins, 388 non-zeros of 12 864: a gather, not a dense contraction -> CUDA cores, no
//     tensor cores), log10, float32
//   * optional fused epilogue: amp[t] = -10 * mean_m(mel[m][t]) in numpy's sequential float32 order
//     (ref:src/aat/tokenizer.py:67), so the boundary kernel does not have to re-read the mel.
//
// Execution shape: persistent CTAs (grid = resident CTAs, three per SM), each looping over tiles of 16
// consecutive frames of one utterance (8 frame pairs x 20 threads = 160 threads).  The raw samples of
// the NEXT tile are fetched with cp.async (LDGSTS, no registers) as soon as pass 1 has consumed the
// current ones, so the global-load latency that dominated the first version of this kernel (ncu: 37 %
// of stall samples on the load->convert dependency, profiles/) hides behind pass 2, the split and the
// mel projection; the filter-bank and log tables are loaded into shared memory once per CTA.
#include "aat_internal.cuh"

namespace aat {

namespace {

constexpr int kFrames = kMelFramesPerTile;   // 16
constexpr int kPairs = kFrames / 2;          // 8
constexpr int kThreads = kPairs * 20;        // 160
constexpr int kRow = 21;                     // padded row (double2 units): 21 is odd -> conflict-free columns
constexpr int kPairStride = 20 * kRow;       // 420 double2; 420 % 8 == 4 keeps neighbouring pairs on distinct banks
constexpr int kPowStride = kBins;            // 201 doubles (odd)
constexpr int kLogTable = 64;                // entries of the log10 table (|r| < 2^-7, degree-8 series)

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

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kBytes>
__device__ __forceinline__ void cp_async(void *dst_smem, const void *src_gmem)
{
    if constexpr (kBytes == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst_smem)), "l"(src_gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "n"(kBytes)
                     : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// log10 of a positive finite double to full double accuracy, table driven:
//   x = 2^e * m, m in [1, 2); i = top 6 mantissa bits; r = m * inv_c[i] - 1 (|r| < 2^-7, one FMA);
//   log10(x) = e * log10(2) + (-log10(inv_c[i])) + log1p(r) / ln(10)
// The table stores inv_c[i] = double(1 / c_i) and -log10 of that ROUNDED value, so the identity is
// exact and the only errors are the final roundings (~1e-16 relative), far below float32 resolution.
__device__ __forceinline__ bool log10_needs_slow_path(double x)
{
    const int hi = (int)(__double_as_longlong(x) >> 32);
    return (unsigned)(hi - 0x00100000) >= 0x7fe00000u; // zero, subnormal, inf, nan, negative
}

// Branch-free: valid for positive normal x; callers patch the rare special values with log10_needs_slow_path.
__device__ __forceinline__ double fast_log10(double x, const double2 *__restrict__ table)
{
    const long long bits = __double_as_longlong(x);
    const int hi = (int)(bits >> 32);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> 14) & (kLogTable - 1);
    const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    const double2 t = table[idx];
    const double r = fma(m, t.x, -1.0);
    // log1p(r)/ln(10) = r * (c1 + r * (c2 + ...)), c_k = (-1)^(k+1) / (k ln 10)
    constexpr double k1 = 0.43429448190325182765112891891661;  //  1/ln10
    constexpr double k2 = -0.21714724095162591382556445945830; // -1/(2 ln10)
    constexpr double k3 = 0.14476482730108394255037630630554;  //  1/(3 ln10)
    constexpr double k4 = -0.10857362047581295691278222972915; // -1/(4 ln10)
    constexpr double k5 = 0.08685889638065036553022578378332;  //  1/(5 ln10)
    constexpr double k6 = -0.07238241365054197127518815315277; // -1/(6 ln10)
    constexpr double k7 = 0.06204206884332168966444698841666;  //  1/(7 ln10)
    constexpr double k8 = -0.05428681023790647845639111486458; // -1/(8 ln10)
    double q = k8;
    q = fma(q, r, k7);
    q = fma(q, r, k6);
    q = fma(q, r, k5);
    q = fma(q, r, k4);
    q = fma(q, r, k3);
    q = fma(q, r, k2);
    q = fma(q, r, k1);
    constexpr double log10_2 = 0.30102999566398119521373889472449;
    return fma((double)e, log10_2, fma(q, r, t.y));
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
    const double2 *log_table;
    const int *mel_row_start;
    const int *mel_bin;
    const double *mel_weight;
    int n_tiles;
    int hop;
    int n_mels;
    int nnz;
    int stage_len; // (kFrames - 1) * hop + 400
    int stage_pad; // stage_len rounded up to 16 bytes worth of samples
};

struct SmemLayout {
    size_t raw, win, tw, logt, mw, mbin, mrow, ex, mel, total;
};

__host__ __device__ inline SmemLayout smem_layout(int stage_pad, int wave_bytes, int n_mels, int nnz)
{
    SmemLayout L{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o += (bytes + 15) & ~size_t(15);
        return at;
    };
    L.ex = take(sizeof(double2) * kPairs * kPairStride); // exchange matrix; later the power spectra (16 x 201 doubles)
    L.raw = take((size_t)stage_pad * wave_bytes);
    L.win = 0; // the window is read through the read-only L1 path (20 coalesced loads at the top of pass 1)
    L.tw = take(sizeof(double2) * 400);
    L.logt = take(sizeof(double2) * kLogTable);
    L.mw = take(sizeof(double) * nnz);
    L.mbin = take(sizeof(int) * n_mels);
    L.mrow = take(sizeof(int) * (n_mels + 1));
    L.mel = L.ex; // float32 mel tile for the amplitude epilogue: overlays the power spectra once they are dead
    L.total = o;
    return L;
}

template <typename WaveT>
__global__ void __launch_bounds__(kThreads, 3) logmel_kernel(const LogmelParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = smem_layout(p.stage_pad, (int)sizeof(WaveT), p.n_mels, p.nnz);
    double2 *s_ex = reinterpret_cast<double2 *>(smem_raw + L.ex);
    double *s_pow = reinterpret_cast<double *>(smem_raw + L.ex);
    WaveT *s_rawbuf = reinterpret_cast<WaveT *>(smem_raw + L.raw);
    const double *__restrict__ g_win = p.window_half;
    double2 *s_tw = reinterpret_cast<double2 *>(smem_raw + L.tw);
    double2 *s_logt = reinterpret_cast<double2 *>(smem_raw + L.logt);
    double *s_mw = reinterpret_cast<double *>(smem_raw + L.mw);
    int *s_mbin = reinterpret_cast<int *>(smem_raw + L.mbin);
    int *s_mrow = reinterpret_cast<int *>(smem_raw + L.mrow);
    float *s_mel = reinterpret_cast<float *>(smem_raw + L.mel);

    const int tid = threadIdx.x;
    constexpr int kVec = 16 / (int)sizeof(WaveT); // samples per 16-byte copy

    // Asynchronous fetch of one tile's raw samples (reflect padding resolved per element,
    // TF:audio_utils.py:769-771); interior, 16-byte aligned tiles move 16 bytes per copy.
    auto prefetch = [&](int tile_id) {
        if (tile_id < p.n_tiles) {
            const int utt = p.tile_utt[tile_id];
            const int64_t n = p.n_samples[utt];
            const int64_t woff = p.wave_off[utt];
            const int64_t g0 = (int64_t)(tile_id - p.tile_first[utt]) * kFrames * p.hop - kNfft / 2;
            const WaveT *wave = reinterpret_cast<const WaveT *>(p.wave) + woff;
            WaveT *dst = s_rawbuf;
            const bool interior = g0 >= 0 && g0 + p.stage_pad <= n;
            const bool aligned = ((woff + g0) % kVec) == 0 && (reinterpret_cast<uintptr_t>(p.wave) & 15) == 0;
            if (interior && aligned) {
                for (int i = tid * kVec; i < p.stage_pad; i += kThreads * kVec) cp_async<16>(dst + i, wave + g0 + i);
            } else {
                for (int i = tid; i < p.stage_len; i += kThreads)
                    cp_async<(int)sizeof(WaveT)>(dst + i, wave + reflect_index(g0 + i, n));
            }
        }
        cp_async_commit();
    };

    prefetch(blockIdx.x);
    for (int i = tid; i < 400; i += kThreads) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < kLogTable; i += kThreads) s_logt[i] = p.log_table[i];
    for (int i = tid; i < p.nnz; i += kThreads) s_mw[i] = p.mel_weight[i];
    for (int i = tid; i < p.n_mels; i += kThreads) s_mbin[i] = p.mel_bin[i];
    for (int i = tid; i <= p.n_mels; i += kThreads) s_mrow[i] = p.mel_row_start[i];

    const int pair = tid / 20;
    const int lane20 = tid - pair * 20;
    double2 *ex = s_ex + pair * kPairStride;
    const int mel_stride = p.n_mels + 1;

    for (int tile_id = blockIdx.x; tile_id < p.n_tiles; tile_id += gridDim.x) {
        const int utt = p.tile_utt[tile_id];
        const int64_t T = 1 + p.n_samples[utt] / p.hop;
        const int64_t f0 = (int64_t)(tile_id - p.tile_first[utt]) * kFrames;
        const int64_t fbase = p.frame_off[utt];

        cp_async_wait<0>(); // this tile's samples have arrived
        __syncthreads();    // ... for every thread; also fences the previous tile's shared-memory reuse

        // ---- pass 1: thread n2 transforms x[20 n1 + n2] over n1, applies W_400^(n2 k1) ----
        {
            double2 v[20];
            const WaveT *wa = s_rawbuf + (2 * pair) * p.hop + lane20;
            const WaveT *wb = wa + p.hop;
#pragma unroll
            for (int n1 = 0; n1 < 20; ++n1) {
                const double w = __ldg(g_win + 20 * n1 + lane20);
                v[n1] = make_double2((double)wa[20 * n1] * w, (double)wb[20 * n1] * w);
            }
            dft20(v);
            ex[lane20] = v[0];
#pragma unroll
            for (int k1 = 1; k1 < 20; ++k1) ex[k1 * kRow + lane20] = cmul(v[k1], s_tw[k1 * 20 + lane20]);
        }
        __syncthreads();
        prefetch(tile_id + gridDim.x); // the raw buffer is free again: the next tile lands during the rest of this one

        // ---- pass 2 + split: thread k1 transforms row k1 over n2 and keeps Z[k1 + 20 k2] in registers ----
        // It owns the bins k = k1 + 20 j (j = 0..9, and j = 10 for k1 = 0).  The mirror Z[400 - k] of those bins
        // sits in row (20 - k1) % 20 at k2 = 19 - j (k1 > 0) or 20 - j (k1 = 0), i.e. always in the UPPER half
        // (k2 >= 10) of the partner row — so only that half goes back through shared memory, and each thread
        // reads 10 mirror values instead of re-reading both operands (-40 % exchange traffic, no division,
        // no bank conflicts).  Round to complex64, power in float64 (TF:audio_utils.py:781,803,808).
        {
            double2 v[20];
            double2 *row = ex + lane20 * kRow;
#pragma unroll
            for (int n2 = 0; n2 < 20; ++n2) v[n2] = row[n2];
            dft20(v);
#pragma unroll
            for (int k2 = 10; k2 < 20; ++k2) row[k2] = v[k2]; // a thread reads and rewrites only its own row
            __syncthreads();
            double pa[11], pb[11];
            const double2 *mir = ex + ((lane20 == 0) ? 0 : 20 - lane20) * kRow;
#pragma unroll
            for (int j = 0; j < 11; ++j) {
                if (j < 10 || lane20 == 0) {
                    const double2 z = v[j];
                    double2 y;
                    if (lane20 == 0)
                        y = (j == 0) ? z : ((j == 10) ? z : mir[20 - j]); // Z[0] and Z[200] mirror themselves
                    else
                        y = mir[19 - j];
                    // X_a = (z + conj y), X_b = (z - conj y) / i   (the 1/2 lives in the window table)
                    const float ar = (float)(z.x + y.x), ai = (float)(z.y - y.y);
                    const float br = (float)(z.y + y.y), bi = (float)(y.x - z.x);
                    pa[j] = (double)ar * (double)ar + (double)ai * (double)ai;
                    pb[j] = (double)br * (double)br + (double)bi * (double)bi;
                }
            }
            __syncthreads(); // every thread has read its mirror values: overlay the matrix with the power spectra
            double *rowa = s_pow + (2 * pair) * kPowStride + lane20;
#pragma unroll
            for (int j = 0; j < 11; ++j) {
                if (j < 10 || lane20 == 0) {
                    rowa[20 * j] = pa[j];
                    rowa[kPowStride + 20 * j] = pb[j];
                }
            }
        }
        __syncthreads();

        // ---- mel projection (banded rows), floor, log10, float32 store ----
        // thread (f, q): frame f = tid % 16, filters q, q + 10, ...  Half-warps share a filter, so the
        // weights are broadcast reads and the 16 frames hit 16 different banks (row stride 201 doubles).
        // Two filters are evaluated together so that two independent dependency chains are in flight.
        constexpr int kGroups = kThreads / kFrames; // 10
        constexpr int kMaxPerThread = (kMaxMels + kGroups - 1) / kGroups;
        constexpr int kBatch = 7; // filters whose logs are evaluated together (7 x 10 groups covers 64..70 filters)
        float outs[kMaxPerThread];
        {
            const int f = tid & (kFrames - 1);
            const int q = tid / kFrames;
            const double *pw = s_pow + f * kPowStride;
            float *dst = p.mel + (size_t)p.n_mels * fbase + f0 + f;
            const bool live = f0 + f < T;
            const int Ti = (int)T;
            auto band = [&](int m) {
                const int j0 = s_mrow[m];
                int n = s_mrow[m + 1] - j0;
                const double *ww = s_mw + j0;
                const double *pp = pw + s_mbin[m];
                double a0 = 0.0, a1 = 0.0;
#pragma unroll 1
                while (n >= 2) { // two independent chains, pointers walk: ~5 instructions per tap
                    a0 = fma(ww[0], pp[0], a0);
                    a1 = fma(ww[1], pp[1], a1);
                    ww += 2, pp += 2, n -= 2;
                }
                if (n) a0 = fma(ww[0], pp[0], a0);
                const double acc = a0 + a1;
                return (acc < 1e-10) ? 1e-10 : acc; // np.maximum(mel_floor, .): NaN propagates
            };
#pragma unroll
            for (int i0 = 0; i0 < kMaxPerThread; i0 += kBatch) {
                if (q + i0 * kGroups < p.n_mels) {
                    // gather the sums of this batch of filters, then take all their logs as one unrolled,
                    // branch-free block: up to seven independent Horner chains in flight per thread
                    double acc[kBatch];
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        const int m = q + (i0 + k) * kGroups;
                        acc[k] = (i0 + k < kMaxPerThread && m < p.n_mels) ? band(m) : 1.0;
                    }
                    double lg[kBatch];
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) lg[k] = fast_log10(acc[k], s_logt);
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        const int m = q + (i0 + k) * kGroups;
                        if (i0 + k < kMaxPerThread && m < p.n_mels) {
                            if (log10_needs_slow_path(acc[k])) lg[k] = log10(acc[k]); // NaN / inf inputs only
                            const float o = (float)lg[k];
                            outs[i0 + k] = o;
                            if (live) dst[(unsigned)(m * Ti)] = o;
                        }
                    }
                }
            }
        }

        // ---- fused amplitude curve: numpy's mean(axis=0) adds the rows in order in float32 ----
        if (p.amp != nullptr) {
            __syncthreads(); // the power spectra are dead: stage the float32 tile over them
            {
                const int f = tid & (kFrames - 1);
                const int q = tid / kFrames;
#pragma unroll
                for (int i = 0; i < kMaxPerThread; ++i) {
                    const int m = q + i * kGroups;
                    if (m < p.n_mels) s_mel[f * mel_stride + m] = outs[i];
                }
            }
            __syncthreads();
            if (tid < kFrames && f0 + tid < T) {
                const float *col = s_mel + tid * mel_stride;
                float acc = col[0];
                for (int m = 1; m < p.n_mels; ++m) acc = __fadd_rn(acc, col[m]);
                const float mean = __fdiv_rn(acc, (float)p.n_mels);
                p.amp[fbase + f0 + tid] = __fmul_rn(-10.0f, mean);
            }
        }
        // the next iteration's first __syncthreads orders these reads before the next overwrite
    }
    cp_async_wait<0>();
}

} // namespace

int logmel_tables_init(aat_ctx *ctx)
{
    // inv_c[i] = double(1 / c_i), c_i = 1 + (i + 0.5) / 128;  y = -log10(inv_c[i]) of the ROUNDED inverse
    std::vector<double2> t(kLogTable);
    for (int i = 0; i < kLogTable; ++i) {
        const long double c = 1.0L + ((long double)i + 0.5L) / (long double)kLogTable;
        const double inv = (double)(1.0L / c);
        t[i] = make_double2(inv, (double)(-log10l((long double)inv)));
    }
    AAT_CUDA_CHECK(cudaMalloc(&ctx->log_table, sizeof(double2) * kLogTable));
    AAT_CUDA_CHECK(cudaMemcpy(ctx->log_table, t.data(), sizeof(double2) * kLogTable, cudaMemcpyHostToDevice));
    return AAT_OK;
}

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
    p.log_table = ctx->log_table;
    p.mel_row_start = ctx->mel.row_start;
    p.mel_bin = ctx->mel.bin;
    p.mel_weight = ctx->mel.weight;
    p.n_tiles = plan->mel_tiles;
    p.hop = ctx->cfg.hop_length;
    p.n_mels = ctx->mel.n_mels;
    p.nnz = ctx->mel.nnz;
    p.stage_len = (kFrames - 1) * p.hop + kNfft;
    const int wave_bytes = wave_dtype == AAT_F32 ? 4 : 8;
    const int vec = 16 / wave_bytes;
    p.stage_pad = (p.stage_len + vec - 1) / vec * vec;
    const size_t smem = smem_layout(p.stage_pad, wave_bytes, p.n_mels, p.nnz).total;
    auto kernel = wave_dtype == AAT_F32 ? logmel_kernel<float> : logmel_kernel<double>;
    AAT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AAT_MAX_SMEM_CARVEOUT(kernel);
    int per_sm = 0;
    AAT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
    AAT_REQUIRE(per_sm >= 1, AAT_ERR_UNSUPPORTED, "aat_logmel: kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    int grid = ctx->num_sms * per_sm;
    if (grid > plan->mel_tiles) grid = plan->mel_tiles;
    ProfileScope prof(ctx, AAT_K_LOGMEL, stream);
    kernel<<<grid, kThreads, smem, stream>>>(p);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat
