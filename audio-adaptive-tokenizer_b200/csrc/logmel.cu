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
// consecutive frames of one utterance (8 frame pairs x 20 threads = 160 threads).  Everything a tile needs to
// know (source offset, output offsets, edge flags) is one 64-byte host-built descriptor; descriptors run two
// tiles ahead and the raw samples one tile ahead, both by cp.async (LDGSTS, no registers), so no thread ever
// waits on a global load inside the loop; tiles beyond a CTA's first three are taken from a global counter, so the
// tail of the kernel does not depend on the tile-count remainder.  The mel projection runs from a host-built banded
// table (MelSchedule): thread (frame, group) walks its own filters with warp-uniform, fully unrolled tap bodies (four
// independent FMA chains), parks the floored sums in its own shared-memory slots and takes their log10 as one
// branch-free, table-driven batch.  Launched with programmatic dependent launch (see aat_internal.cuh).
#include "aat_internal.cuh"

namespace aat {

namespace {

constexpr int kFrames = kMelFramesPerTile;   // 16
constexpr int kPairs = kFrames / 2;          // 8
constexpr int kThreads = kPairs * 20;        // 160
constexpr int kRow = 21;                     // padded row (double2 units): 21 is odd -> conflict-free columns
constexpr int kPairStride = 20 * kRow;       // 420 double2; 420 % 8 == 4 keeps neighbouring pairs on distinct banks
constexpr int kPowStride = kBins;            // 201 doubles (odd)
constexpr int kLogTable = 32;                // entries of the log10 table (|r| < 2^-6, degree-8 series: |error| < 3e-18)
constexpr int kLogCopies = 8;                // shared-memory replicas: lane l reads copy l & 7, so the eight lanes of a
                                             // quarter-warp hit eight different 16-byte bank groups whatever their indices

// FP64 literals cost two UMOVs per use as immediates (and the compiler folds __constant__ initialisers back
// into immediates); as kernel parameters they are constant-bank operands of DFMA/DADD, i.e. free.
struct LogmelConsts {
    double c1, c2, s1, s2; // cos(2 pi/5), cos(4 pi/5), sin(2 pi/5), sin(4 pi/5)
    double k[8];           // log1p(r)/ln(10) = r (k[0] + r (k[1] + ...)), k[i] = (-1)^i / ((i + 1) ln 10)
    double log10_2;
    double int_magic;      // 2^52 + 2^31
};
static const LogmelConsts kLogmelConsts = {
    0.30901699437494742410229341718282,  -0.80901699437494742410229341718282,
    0.95105651629515357211643933337938,  0.58778525229247312916870595463907,
    {0.43429448190325182765112891891661, -0.21714724095162591382556445945830, 0.14476482730108394255037630630554,
     -0.10857362047581295691278222972915, 0.08685889638065036553022578378332, -0.07238241365054197127518815315277,
     0.06204206884332168966444698841666, -0.05428681023790647845639111486458},
    0.30102999566398119521373889472449,
    4503601774854144.0,
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// double(float(x)) without the two conversions: they run on the 16-lane XU path (8.3 + 5.9 issue cycles per scheduler,
// 23 + 19 cycles of latency, profiles/r1_ubench_fp64.txt) while the FP64 pipe has room.  Veltkamp's splitting with
// C = 2^29 + 1 returns x rounded to nearest on 24 significant bits (three dependent FP64 operations).  It differs from
// the conversion pair only for exact ties (2^-29 of the values, where it may round away from even) and outside
// float32's normal range (|x| < 1.2e-38: the squared magnitude is then below the mel floor by 60 orders of magnitude).
__device__ __forceinline__ double round_to_float(double x)
{
#ifdef AAT_LOGMEL_F2F
    return (double)(float)x;
#else
    const double g = __dmul_rn(x, 536870913.0);
    return __dadd_rn(g, __dsub_rn(x, g));
#endif
}

// 20-point forward DFT in registers, Good-Thomas 4 x 5:
//   n = (5 n1 + 4 n2) mod 20,  k = (5 k1 + 16 k2) mod 20,  X[k] = sum x[n] W4^(n1 k1) W5^(n2 k2)
__device__ __forceinline__ void dft20(double2 (&v)[20], const LogmelConsts &K)
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
    const double c1 = K.c1, c2 = K.c2, s1 = K.s1, s2 = K.s2;
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
//   x = 2^e * m, m in [1, 2); i = top 5 mantissa bits; r = m * inv_c[i] - 1 (|r| < 2^-6, one FMA);
//   log10(x) = e * log10(2) + (-log10(inv_c[i])) + log1p(r) / ln(10)
// The table stores inv_c[i] = double(1 / c_i) and -log10 of that ROUNDED value, so the identity is
// exact and the only errors are the final roundings (~1e-16 relative), far below float32 resolution.
__device__ __forceinline__ bool log10_needs_slow_path(double x)
{
    const int hi = (int)(__double_as_longlong(x) >> 32);
    return (unsigned)(hi - 0x00100000) >= 0x7fe00000u; // zero, subnormal, inf, nan, negative
}

// Branch-free: valid for positive normal x; callers patch the rare special values with log10_needs_slow_path.
__device__ __forceinline__ double fast_log10(double x, const double2 *__restrict__ table, const LogmelConsts &K)
{
    const long long bits = __double_as_longlong(x);
    const int hi = (int)(bits >> 32);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> 15) & (kLogTable - 1);
    const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    const double2 t = table[idx * kLogCopies + (threadIdx.x & (kLogCopies - 1))];
    const double r = fma(m, t.x, -1.0);
    double q = K.k[7];
    q = fma(q, r, K.k[6]);
    q = fma(q, r, K.k[5]);
    q = fma(q, r, K.k[4]);
    q = fma(q, r, K.k[3]);
    q = fma(q, r, K.k[2]);
    q = fma(q, r, K.k[1]);
    q = fma(q, r, K.k[0]);
    // double(e) without I2F: 2^52 + 2^31 + e is exact in the low word of a double
    const double ed = __hiloint2double(0x43300000, e ^ (int)0x80000000) - K.int_magic;
    return fma(ed, K.log10_2, fma(q, r, t.y));
}

// zero, subnormal, inf, nan, negative: the library routine, kept out of line (it is ~250 instructions and
// would otherwise be inlined at every call site of the unrolled log batch)
__device__ __noinline__ double slow_log10(double x) { return log10(x); }

struct LogmelParams {
    const double *znorm; // optional [2 * n_utts] (mean, population variance): fused z-score of the samples
    const void *wave;
    float *mel;
    float *amp;
    const MelTile *tiles;
    const double *window_half;
    const double2 *twiddle;
    const double2 *log_table;
    const uint16_t *chunk_desc;
    const double *mel_weight;
    int n_tiles;
    int hop;
    int n_mels;
    int n_weights;
    int n_desc;    // uint16 entries of chunk_desc
    int stage_len; // (kFrames - 1) * hop + 400
    int stage_pad; // stage_len rounded up to 16 bytes worth of samples
    int *sched;    // [2] dynamic tile counter, exit counter (both left at zero by the last CTA to exit)
    LogmelConsts K;
};

constexpr int kTwiddles = 7 * 20; // rows k1 = 1, 2, 3, 4, 5, 10, 15
constexpr int kTileRing = 3; // descriptors: current tile, the tile whose samples are being fetched, the one after

// Raw-sample staging for hop 160: frame pairs start 320 samples apart, a multiple of the 32 banks, so the lanes of
// a warp that belong to the next pair would collide with the first pair's (2-way conflict on all 40 sample loads of
// pass 1).  A gap after every 320 samples shifts each pair by 20 banks: then bank = thread index, conflict-free.
constexpr int kRawBlock = 320;
template <typename WaveT>
__host__ __device__ constexpr int raw_gap() { return sizeof(WaveT) == 4 ? 20 : 4; } // (20 - 320) mod 32 words / mod 16 double words
__host__ __device__ inline int raw_elems(int stage_pad, int gap) { return stage_pad + gap * ((stage_pad - 1) / kRawBlock); }

struct SmemLayout {
    size_t ex, raw, tw, logt, mw, fdesc, tiles, next, mel, acc, total;
};

__host__ __device__ inline SmemLayout smem_layout(int raw_elems, int wave_bytes, int n_desc, int n_weights)
{
    SmemLayout L{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o += (bytes + 15) & ~size_t(15);
        return at;
    };
    // exchange matrix of the two FFT passes; afterwards the power spectra (16 x 201 doubles) and, behind them, the
    // float32 mel tile of the amplitude epilogue
    L.ex = take(sizeof(double2) * kPairs * kPairStride);
    L.mel = L.ex + sizeof(double) * kFrames * kPowStride;
    L.acc = (L.mel + sizeof(float) * kFrames * (kMaxMels + 1) + 15) & ~size_t(15); // per-thread filter sums [13][160]
    L.raw = take((size_t)raw_elems * wave_bytes);
    L.tw = take(sizeof(double2) * kTwiddles);
    L.logt = take(sizeof(double2) * kLogTable * kLogCopies);
    L.mw = take(sizeof(double) * n_weights);
    L.fdesc = take(sizeof(uint16_t) * n_desc);
    L.tiles = take(sizeof(MelTile) * kTileRing);
    L.next = take(sizeof(int)); // the tile id thread 0 grabbed during the current tile
    L.total = o;
    return L;
}
static_assert(sizeof(double) * kFrames * kPowStride + sizeof(float) * kFrames * (kMaxMels + 1) + 16 +
                      sizeof(double) * ((kMaxMels + kMelGroups - 1) / kMelGroups) * kThreads <=
                  sizeof(double2) * kPairs * kPairStride,
              "power spectra + float32 mel tile + staged filter sums must fit in the exchange matrix");

// First / last tiles of an utterance (and unaligned ones): element-wise copies with np.pad(mode="reflect")
// index arithmetic (TF:audio_utils.py:769-771).  Rare, so out of line: the 64-bit modulo is bulky.
template <typename WaveT>
__device__ __noinline__ void fetch_edge_tile(WaveT *dst, const WaveT *utt, int64_t g0, int64_t n, int stage_len, int gap,
                                             int tid)
{
#pragma unroll 1
    for (int i = tid; i < stage_len; i += kThreads)
        cp_async<(int)sizeof(WaveT)>(dst + i + gap * (i / kRawBlock), utt + reflect_index(g0 + i, n));
}

// Fused z-score (x - mean) / (std + 1e-6) of the call sites (ref:src/aat/training/collate.py:135-152,
// ref:scripts/audio_tokenization_melspec.py:40) in float64, applied where a sample is widened for the transform, so a
// normalised copy of the waveform is never written or re-read: the Znorm functor of aat_internal.cuh (correctly rounded
// quotient in three FP64 operations), the very one aat_normalize applies, so fused and separate passes agree bit for bit.
AAT_TIMELINE_STORAGE(logmel)
template <typename WaveT, bool kHop160, bool kZnorm>
__global__ void __launch_bounds__(kThreads, 3) logmel_kernel(const LogmelParams p)
{
    AAT_TIMELINE_SCOPE(logmel);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kGap = kHop160 ? raw_gap<WaveT>() : 0;
    const SmemLayout L = smem_layout(raw_elems(p.stage_pad, kGap), (int)sizeof(WaveT), p.n_desc, p.n_weights);
    double2 *s_ex = reinterpret_cast<double2 *>(smem_raw + L.ex);
    double *s_pow = reinterpret_cast<double *>(smem_raw + L.ex);
    WaveT *s_rawbuf = reinterpret_cast<WaveT *>(smem_raw + L.raw);
    const double *__restrict__ g_win = p.window_half;
    double2 *s_tw = reinterpret_cast<double2 *>(smem_raw + L.tw);
    double2 *s_logt = reinterpret_cast<double2 *>(smem_raw + L.logt);
    double *s_mw = reinterpret_cast<double *>(smem_raw + L.mw);
    uint16_t *s_cdesc = reinterpret_cast<uint16_t *>(smem_raw + L.fdesc);
    MelTile *s_tiles = reinterpret_cast<MelTile *>(smem_raw + L.tiles);
    int *s_next = reinterpret_cast<int *>(smem_raw + L.next);
    float *s_mel = reinterpret_cast<float *>(smem_raw + L.mel);
    double *s_acc = reinterpret_cast<double *>(smem_raw + L.acc);

    const int tid = threadIdx.x;
    constexpr int kVec = 16 / (int)sizeof(WaveT); // samples per 16-byte copy
    const bool wave_aligned = (reinterpret_cast<uintptr_t>(p.wave) & 15) == 0;

    // Descriptor of tile `tile_id` -> ring slot (asynchronous; four threads move 16 bytes each).  Past the end
    // the slot is marked "no tile".  Joins the cp.async group of the sample fetch that follows it.
    auto fetch_desc = [&](int tile_id, int slot) {
        if (tid < 4) {
            if (tile_id < p.n_tiles)
                cp_async<16>(reinterpret_cast<unsigned char *>(s_tiles + slot) + 16 * tid,
                             reinterpret_cast<const unsigned char *>(p.tiles + tile_id) + 16 * tid);
            else if (tid == 3)
                s_tiles[slot].valid = 0;
        }
    };
    // Asynchronous fetch of one tile's raw samples (reflect padding resolved per element,
    // TF:audio_utils.py:769-771); interior, 16-byte aligned tiles move 16 bytes per copy.
    auto fetch_samples = [&](const MelTile &d) {
        if (d.valid > 0) {
            const WaveT *wave = reinterpret_cast<const WaveT *>(p.wave);
            const int64_t src = d.src;
            if (d.interior && wave_aligned && (src & (kVec - 1)) == 0) {
                const WaveT *from = wave + src;
                if constexpr (kHop160) {
                    // 2 800 samples: the trip count and the gap of every copy are known at compile time
                    constexpr int kStage = (kFrames - 1) * 160 + kNfft;
                    constexpr int kPerBlock = kRawBlock / kVec; // 16-byte copies per 320-sample block
                    static_assert(kStage % kVec == 0 && kThreads % kPerBlock == 0, "staging geometry");
                    const int at = tid * kVec + kGap * (tid / kPerBlock);
#pragma unroll
                    for (int k = 0; k * kThreads * kVec < kStage; ++k)
                        if ((k + 1) * kThreads * kVec <= kStage || (tid + k * kThreads) * kVec < kStage)
                            cp_async<16>(s_rawbuf + at + k * (kThreads * kVec + kGap * (kThreads / kPerBlock)),
                                         from + (tid + k * kThreads) * kVec);
                } else {
                    for (int i = tid * kVec; i < p.stage_pad; i += kThreads * kVec) cp_async<16>(s_rawbuf + i, from + i);
                }
            } else {
                fetch_edge_tile<WaveT>(s_rawbuf, wave + d.wave_off, src - d.wave_off, d.n, p.stage_len, kGap, tid);
            }
        }
        cp_async_commit();
    };

    // ---- prologue: descriptors of the first two tiles, samples of the first, the per-CTA tables ----
    fetch_desc(blockIdx.x, 0);
    fetch_desc(blockIdx.x + gridDim.x, 1);
    cp_async_commit();
    for (int i = tid; i < kTwiddles; i += kThreads) s_tw[i] = p.twiddle[i];
    for (int i = tid; i < kLogTable * kLogCopies; i += kThreads) s_logt[i] = p.log_table[i / kLogCopies];
    for (int i = tid; i < p.n_weights; i += kThreads) s_mw[i] = p.mel_weight[i];
    for (int i = tid; i < p.n_desc; i += kThreads) s_cdesc[i] = p.chunk_desc[i];
    cp_async_wait<0>();
    __syncthreads();
    // everything above reads tables that no kernel writes; the samples may come from the previous kernel
    pdl_wait();
    pdl_launch_dependents();
    fetch_samples(s_tiles[0]);

    const int pair = tid / 20;
    const int lane20 = tid - pair * 20;
    double2 *ex = s_ex + pair * kPairStride;
    const int mel_stride = p.n_mels + 1;
    const int f = tid & (kFrames - 1); // mel / log phases: frame of the tile
    const int q = tid / kFrames;       // ... and thread group (filters q, q + 10, ...)
    const int n_mine = q < p.n_mels ? (p.n_mels - q + kMelGroups - 1) / kMelGroups : 0; // filters of this thread

    // Tile schedule: the first three tiles of a CTA are static (blockIdx + k * grid); every later one is grabbed from a
    // global counter three tiles ahead of its use (descriptor two ahead, samples one ahead), so the tile-count
    // remainder (6404 tiles over 444 CTAs = 14.4 each) and slow CTAs even out instead of setting the kernel's tail.
    // Thread 0 issues the atomic right after a tile's first barrier and parks the result in shared memory before
    // the second one: its latency hides behind pass 1.
    const int G = (int)gridDim.x;
    int tile_id = blockIdx.x, next_id = tile_id + G, after_id = next_id + G;

    int slot = 0; // ring slot of the current tile
    while (tile_id < p.n_tiles) {
        const int slot_next = slot + 1 == kTileRing ? 0 : slot + 1;
        const int slot_after = slot_next + 1 == kTileRing ? 0 : slot_next + 1;

        cp_async_wait<0>(); // this tile's samples and the next tile's descriptor have arrived
        __syncthreads();    // ... for every thread; also fences the previous tile's shared-memory reuse
        int grabbed = 0;
        if (tid == 0) grabbed = atomicAdd(p.sched, 1);

        // ---- pass 1: thread n2 transforms x[20 n1 + n2] over n1, applies W_400^(n2 k1) ----
        {
            double2 v[20];
            Znorm zn{0.0, 1.0, 1.0};
            if (kZnorm) zn = Znorm::from_stats(p.znorm, s_tiles[slot].utt);
            const WaveT *wa = s_rawbuf + pair * (kHop160 ? kRawBlock + kGap : 2 * p.hop) + lane20;
            const int hop = kHop160 ? 160 : p.hop;
#pragma unroll
            for (int n1 = 0; n1 < 20; ++n1) {
                const double w = __ldg(g_win + 20 * n1 + lane20);
                // frame a: samples 20 n1 + lane20 of the pair's block; frame b: one hop later.  With the gapped layout
                // the offsets that cross into the next 320-sample block are known at compile time.
                const int ia = 20 * n1 + (kHop160 && 20 * n1 >= kRawBlock ? kGap : 0);
                const int ib = hop + 20 * n1 + (kHop160 && 160 + 20 * n1 >= kRawBlock ? kGap : 0);
                if (kZnorm)
                    v[n1] = make_double2(zn((double)wa[ia]) * w, zn((double)wa[ib]) * w);
                else
                    v[n1] = make_double2((double)wa[ia] * w, (double)wa[ib] * w);
            }
            dft20(v, p.K);
            // W_400^(n2 k1) for k1 = 5 j + i is (W^(5 j n2)) (W^(i n2)): seven table rows (k1 = 1..4, 5, 10, 15) and twelve
            // complex products instead of nineteen 16-byte loads per thread — the shared-memory pipe is this kernel's
            // busiest unit, the FP64 pipe is not.  One extra rounding per derived twiddle (1e-16 relative).
            ex[lane20] = v[0];
            double2 w[5];
#pragma unroll
            for (int i = 1; i < 5; ++i) {
                w[i] = s_tw[(i - 1) * 20 + lane20];
                ex[i * kRow + lane20] = cmul(v[i], w[i]);
            }
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                const double2 a = s_tw[(3 + j) * 20 + lane20];
                ex[(5 * j) * kRow + lane20] = cmul(v[5 * j], a);
#pragma unroll
                for (int i = 1; i < 5; ++i) ex[(5 * j + i) * kRow + lane20] = cmul(v[5 * j + i], cmul(a, w[i]));
            }
        }
        if (tid == 0) *s_next = 3 * G + grabbed;
        __syncthreads();
        const int grabbed_id = *s_next; // this CTA's tile after `after_id`
        // the raw buffer is free again: the next tile lands during the rest of this one
        fetch_desc(after_id, slot_after);
        fetch_samples(s_tiles[slot_next]);

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
            dft20(v, p.K);
#pragma unroll
            for (int k2 = 10; k2 < 20; ++k2) row[k2] = v[k2]; // a thread reads and rewrites only its own row
            __syncthreads();
            double pa[11], pb[11];
            const double2 *mir = ex + ((lane20 == 0) ? 0 : 20 - lane20) * kRow;
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const double2 z = v[j];
                double2 y;
                if (lane20 == 0)
                    y = (j == 0) ? z : mir[20 - j]; // Z[0] mirrors itself
                else
                    y = mir[19 - j];
                // X_a = (z + conj y), X_b = (z - conj y) / i   (the 1/2 lives in the window table)
                const double ar = round_to_float(z.x + y.x), ai = round_to_float(z.y - y.y);
                const double br = round_to_float(z.y + y.y), bi = round_to_float(y.x - z.x);
                pa[j] = ar * ar + ai * ai;
                pb[j] = br * br + bi * bi;
            }
            if (lane20 == 0) { // bin 200 = Z[200] of row 0 mirrors itself: both spectra are real there
                const double ar = round_to_float(v[10].x + v[10].x), br = round_to_float(v[10].y + v[10].y);
                pa[10] = ar * ar;
                pb[10] = br * br;
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

        // ---- mel projection, floor, log10, float32: thread (f, q) finishes filters q, q + 10, ... of frame f ----
        // A half-warp is 16 frames of one filter: the weights are broadcast reads and the power values hit 16
        // different banks (row stride 201 doubles).  The two half-warps of a warp own neighbouring filters padded to one
        // length, so the switch on the tap count does not diverge.  The sums stay in registers: no second pass, no
        // barrier between projection and log.
        const MelTile &cur = s_tiles[slot];
        {
            constexpr int kMaxPerThread = (kMaxMels + kMelGroups - 1) / kMelGroups;
            constexpr int kBatch = 7; // filters whose logs are evaluated together (7 x 10 groups covers 64..70 filters)
            const double *pw = s_pow + f * kPowStride;
            float *dst = p.mel + cur.mel_off + f;
            const bool live = f < cur.valid;
            const size_t T = (size_t)cur.T;
            float *smel = s_mel + f * mel_stride;
            // projection: ONE rolled loop over the chunk stream of this thread's group (four taps per chunk = four
            // independent FMA chains, explicit fma / _rn so that the sums do not depend on the compiler's mood).  The
            // operands of chunk c + 1 are fetched before chunk c is added up; a chunk that ends a filter reduces the
            // chains ((a0 + a1) + (a2 + a3)), floors the sum and parks it in the thread's own shared-memory slot for
            // the unrolled log batch below.  Zero weights pad a filter to whole chunks: they read this frame's own
            // power values, so they add exact zeros (or keep a NaN frame NaN).
            {
                double *my_acc = s_acc + tid;
                const uint32_t first_chunk = s_cdesc[3 * q], n_chunks = s_cdesc[3 * q + 1];
                const uint16_t *cd = s_cdesc + kMelDescHeader + first_chunk;
                const double2 *w2 = reinterpret_cast<const double2 *>(s_mw) + 2 * first_chunk;
                const double *pp = pw + s_cdesc[3 * q + 2];
                double2 wa = w2[0], wb = w2[1];
                double p0 = pp[0], p1 = pp[1], p2 = pp[2], p3 = pp[3];
                uint32_t d = cd[0];
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
                for (uint32_t c = 0; c < n_chunks; ++c) {
                    w2 += 2;
                    ++cd;
                    const double2 nwa = w2[0], nwb = w2[1];
                    const double *pn = pw + (d & 0xffu);
                    const double n0 = pn[0], n1 = pn[1], n2 = pn[2], n3 = pn[3];
                    const uint32_t dn = cd[0];
                    a0 = fma(wa.x, p0, a0);
                    a1 = fma(wa.y, p1, a1);
                    a2 = fma(wb.x, p2, a2);
                    a3 = fma(wb.y, p3, a3);
                    if (d & kMelDescLast) {
                        const double acc = __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));
                        *my_acc = (acc < 1e-10) ? 1e-10 : acc; // np.maximum(mel_floor, .): NaN propagates
                        my_acc += kThreads;
                        a0 = a1 = a2 = a3 = 0.0;
                    }
                    wa = nwa, wb = nwb;
                    p0 = n0, p1 = n1, p2 = n2, p3 = n3;
                    d = dn;
                }
            }
            const double *my_acc = s_acc + tid;
#pragma unroll
            for (int i0 = 0; i0 < kMaxPerThread; i0 += kBatch) {
                if (i0 < n_mine) {
                    // gather the sums of this batch of filters, then take all their logs as one unrolled,
                    // branch-free block: up to seven independent Horner chains in flight per thread
                    double acc[kBatch];
#pragma unroll
                    for (int k = 0; k < kBatch; ++k)
                        acc[k] = (i0 + k < kMaxPerThread && i0 + k < n_mine) ? my_acc[(i0 + k) * kThreads] : 1.0;
                    double lg[kBatch];
                    unsigned worst = 0; // after the floor a value is >= 1e-10, +inf or NaN: one test for the whole batch
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        lg[k] = fast_log10(acc[k], s_logt, p.K);
                        worst = max(worst, (unsigned)__double2hiint(acc[k]));
                    }
                    if (worst >= 0x7ff00000u) {
#pragma unroll
                        for (int k = 0; k < kBatch; ++k) // unrolled: a runtime index would push acc/lg into local memory
                            if (log10_needs_slow_path(acc[k])) lg[k] = slow_log10(acc[k]); // NaN / inf inputs only
                    }
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        const int m = q + (i0 + k) * kMelGroups;
                        if (i0 + k < kMaxPerThread && i0 + k < n_mine) {
                            const float o = (float)lg[k];
                            smel[m] = o; // float32 tile of the amplitude epilogue (behind the power spectra)
                            if (live) dst[(size_t)m * T] = o;
                        }
                    }
                }
            }
        }

        // ---- fused amplitude curve: numpy's mean(axis=0) adds the rows in order in float32 ----
        if (p.amp != nullptr) {
            __syncthreads();
            if (tid < kFrames && tid < cur.valid) {
                // The additions are one dependent chain, the loads are not: batches of 21 loads are in flight together
                // (64 filters = the first term + three batches); a rolled loop paid a shared-memory round trip every
                // four terms.  -0.8 us per config-2 launch (profiles/r2_logmel_ab.txt).
                const float *col = s_mel + tid * mel_stride;
                const int n = p.n_mels;
                float acc = col[0];
                int m = 1;
#pragma unroll 1
                for (; m + 21 <= n; m += 21) {
                    float v[21];
#pragma unroll
                    for (int k = 0; k < 21; ++k) v[k] = col[m + k];
#pragma unroll
                    for (int k = 0; k < 21; ++k) acc = __fadd_rn(acc, v[k]);
                }
                for (; m < n; ++m) acc = __fadd_rn(acc, col[m]);
                const float mean = __fdiv_rn(acc, (float)p.n_mels);
                p.amp[cur.amp_off + tid] = __fmul_rn(-10.0f, mean);
            }
        }
        // the next iteration's first __syncthreads orders these reads before the next overwrite
        slot = slot_next;
        tile_id = next_id, next_id = after_id, after_id = grabbed_id;
    }
    cp_async_wait<0>();
    // the last CTA to leave resets the tile scheduler for the next launch (graph replays included)
    if (tid == 0 && atomicAdd(p.sched + 1, 1) == G - 1) {
        p.sched[0] = 0;
        p.sched[1] = 0;
    }
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

template <typename WaveT>
static auto pick_logmel_kernel(bool hop160, bool znorm)
{
    if (hop160) return znorm ? logmel_kernel<WaveT, true, true> : logmel_kernel<WaveT, true, false>;
    return znorm ? logmel_kernel<WaveT, false, true> : logmel_kernel<WaveT, false, false>;
}

int launch_logmel(aat_ctx *ctx, const aat_plan *plan, const void *wave, int wave_dtype, const double *znorm_stats,
                  float *mel, float *amp, cudaStream_t stream)
{
    AAT_REQUIRE(wave_dtype == AAT_F32 || wave_dtype == AAT_F64, AAT_ERR_UNSUPPORTED,
                "aat_logmel: waveform dtype must be AAT_F32 or AAT_F64 (got %d)", wave_dtype);
    if (plan->mel_tiles == 0) return AAT_OK;
    LogmelParams p{};
    p.znorm = znorm_stats;
    p.wave = wave;
    p.mel = mel;
    p.amp = amp;
    p.tiles = plan->d_mel_tile;
    p.window_half = ctx->window_half;
    p.twiddle = ctx->twiddle;
    p.log_table = ctx->log_table;
    p.chunk_desc = ctx->mel.chunk_desc;
    p.mel_weight = ctx->mel.weight;
    p.n_tiles = plan->mel_tiles;
    p.hop = ctx->cfg.hop_length;
    p.n_mels = ctx->mel.n_mels;
    p.n_weights = ctx->mel.n_weights;
    p.n_desc = ctx->mel.n_desc;
    p.K = kLogmelConsts;
    p.sched = plan->d_mel_sched;
    p.stage_len = (kFrames - 1) * p.hop + kNfft;
    const int wave_bytes = wave_dtype == AAT_F32 ? 4 : 8;
    const int vec = 16 / wave_bytes;
    p.stage_pad = (p.stage_len + vec - 1) / vec * vec;
    const bool hop160 = p.hop == 160;
    const int gap = hop160 ? (wave_dtype == AAT_F32 ? raw_gap<float>() : raw_gap<double>()) : 0;
    const size_t smem = smem_layout(raw_elems(p.stage_pad, gap), wave_bytes, p.n_desc, p.n_weights).total;
    const bool znorm = znorm_stats != nullptr;
    auto kernel = wave_dtype == AAT_F32 ? pick_logmel_kernel<float>(hop160, znorm) : pick_logmel_kernel<double>(hop160, znorm);
    int per_sm = 0;
    AAT_CUDA_CHECK(prepare_kernel(ctx, kernel, kThreads, smem, &per_sm));
    AAT_REQUIRE(per_sm >= 1, AAT_ERR_UNSUPPORTED, "aat_logmel: kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    int grid = ctx->num_sms * per_sm;
    // Tried and rejected (gpurun r2_b3 / r2_b4): leaving one CTA slot per utterance free on long streams, so that the
    // previous batch's boundary scan could run beside this kernel in the pipelined schedule (aat_b200/pipeline.py):
    // the free slots do not end up one per SM, the scan's 256-thread CTAs did not fit them, and nothing overlapped
    // (config 4: 2.20 -> 2.23 ms per step at depth 2).  A third batch in flight does hide the scan (2.07 ms).
    if (grid > plan->mel_tiles) grid = plan->mel_tiles;
    ProfileScope prof(ctx, AAT_K_LOGMEL, stream);
    AAT_CUDA_CHECK(launch_pdl(kernel, dim3(grid), dim3(kThreads), smem, stream, p));
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat

AAT_TIMELINE_EXPORT(logmel, aat::)
