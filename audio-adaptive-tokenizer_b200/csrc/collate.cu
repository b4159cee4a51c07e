// Callers on either side of the hot path — HBM-bound byte movers.
//
// N2  aat_normalize: the per-utterance normalisations that feed the path
//       z-score   (x - mean) / (std + 1e-6)        ref:scripts/audio_tokenization_melspec.py:40,
//                                                   ref:src/aat/training/collate.py:135,138,152
//       wav2vec2  (x - mean) / sqrt(var + 1e-7)    ref:src/aat/training/collate.py:301 ->
//                                                   TF:models/wav2vec2/feature_extraction_wav2vec2.py:78-95
//     Statistics in float64: every 4096-sample chunk yields (n, mean, M2) with a two-pass reduction inside
//     the CTA; an utterance's chunks are merged in order with Chan's formula (robust against a DC offset,
//     deterministic, one pass over HBM); a second kernel applies the affine map.
// N1  the collator's ragged -> padded layout (ref:src/aat/training/collate.py:242-253, 291-346), which the
//     reference builds with a double Python loop ("todo vectorize", :248):
//       aat_pad_segment_boarders   cumulative segment ends -> [B, S_max] int64, zero padded, + mask
//       aat_scatter_segments       waveform slices  -> [B, S_max, max_frames] float32 + mask
//       aat_scatter_mel_segments   log-mel slices   -> [B, S_max, n_mels, max_items] float32
//     One CTA per (utterance, segment) row; 128-bit stores; zero fill of the padding in the same pass.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "aat_internal.cuh"

namespace aat {

namespace {

constexpr int kStatThreads = 256;
constexpr int kStatPer = 16;                         // samples per thread
constexpr int kStatChunk = kStatThreads * kStatPer;  // 4096 samples per CTA
static_assert(kStatChunk == kNormChunk, "the plan's chunk tables are built for kNormChunk samples");

struct Moments { // (count, mean, sum of squared deviations) of a set of samples
    double n, mean, m2;
};
// Chan et al.: moments of the union of two disjoint sets (robust against a DC offset; a fixed merge tree keeps the
// result deterministic)
__device__ __forceinline__ Moments merge(const Moments &a, const Moments &b)
{
    const double tot = a.n + b.n;
    if (tot == 0.0) return Moments{0.0, 0.0, 0.0};
    const double delta = b.mean - a.mean;
    const double w = b.n / tot;
    return Moments{tot, fma(delta, w, a.mean), a.m2 + b.m2 + delta * delta * (a.n * w)};
}
__device__ __forceinline__ Moments shfl_xor(const Moments &v, int d)
{
    return Moments{__shfl_xor_sync(0xffffffffu, v.n, d), __shfl_xor_sync(0xffffffffu, v.mean, d),
                   __shfl_xor_sync(0xffffffffu, v.m2, d)};
}

// 16 samples of thread `t` of a chunk.  Slot s of thread t is sample slot_index<WaveT>(s, t) of the chunk: 16-byte
// vector q = s / kVec of the thread is vector q * 256 + t of the chunk, so every load / store instruction of a warp
// covers one contiguous 512-byte range.  Whole, aligned chunks move by 16-byte accesses, the others element by element.
template <typename T>
__device__ __forceinline__ int slot_index(int s, int t)
{
    constexpr int kVec = 16 / (int)sizeof(T);
    return ((s / kVec) * kStatThreads + t) * kVec + (s % kVec);
}
// The samples stay in their own type in registers (16 floats are 16 registers, 16 doubles 32): with the converted
// copies the statistics kernel needed 56 registers, four CTAs per SM, and left the memory system idle during its two
// block reductions (33 % DRAM utilisation, ncu r2_prof_collate); they are widened where they are used.
template <typename WaveT>
__device__ __forceinline__ void load16(const WaveT *src, int64_t len, int t, WaveT (&v)[kStatPer])
{
    constexpr int kVec = 16 / (int)sizeof(WaveT);
    if (len == kStatChunk && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < kStatPer / kVec; ++q) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(src) + q * kStatThreads + t);
            const WaveT *e = reinterpret_cast<const WaveT *>(&raw);
#pragma unroll
            for (int k = 0; k < kVec; ++k) v[q * kVec + k] = e[k];
        }
    } else {
#pragma unroll
        for (int s = 0; s < kStatPer; ++s) {
            const int i = slot_index<WaveT>(s, t);
            v[s] = (i < len) ? src[i] : (WaveT)0;
        }
    }
}

// sum over the CTA, returned to every thread: shuffle tree per warp, one shared-memory hop, fixed order (deterministic)
__device__ __forceinline__ double block_sum(double v, double *s_red)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kStatThreads / 32; ++w) t += s_red[w];
    return t;
}

// grid = total chunks (chunk_utt / chunk_first tables as for the mel tiles); partial[chunk] = (n, mean, M2).
// Two-pass moments of the chunk with the samples held in registers: sum -> mean -> sum of squared deviations.  Plain
// additions only (merging per-thread moments with Chan's formula cost a float64 division per shuffle step: 50 us per
// config-2 batch against the 10 us the 65 MB take at the HBM peak; gpurun r2_t7).  One CTA per chunk: persistent CTAs
// with the next chunk's samples in flight were tried and are slower (32 vs 23 us: fewer loads in flight in total).
template <typename WaveT>
__global__ void __launch_bounds__(kStatThreads, sizeof(WaveT) == 4 ? 8 : 5)
wave_chunk_stats_kernel(const WaveT *wave, const int64_t *n_samples, const int64_t *wave_off, const int32_t *chunk_utt,
                        const int32_t *chunk_first, double *partial)
{
    __shared__ double s_red[2][kStatThreads / 32];
    const int utt = chunk_utt[blockIdx.x];
    const int64_t c = blockIdx.x - chunk_first[utt];
    const int64_t n = n_samples[utt];
    const int64_t j0 = c * kStatChunk;
    const int64_t len = (n - j0 < kStatChunk) ? n - j0 : kStatChunk;
    WaveT v[kStatPer];
    load16(wave + wave_off[utt] + j0, len, threadIdx.x, v);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kStatPer; ++k) s += (double)v[k]; // slots past the end hold 0
    const double mean = block_sum(s, s_red[0]) / (double)len;
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < kStatPer; ++k) {
        const double d = (double)v[k] - mean;
        q += (len == kStatChunk || slot_index<WaveT>(k, threadIdx.x) < len) ? d * d : 0.0;
    }
    const double m2 = block_sum(q, s_red[1]);
    if (threadIdx.x == 0) {
        partial[3 * (size_t)blockIdx.x + 0] = (double)len;
        partial[3 * (size_t)blockIdx.x + 1] = mean;
        partial[3 * (size_t)blockIdx.x + 2] = m2;
    }
}

// one WARP per utterance merges its chunk moments: lane l takes chunks l, l + 32, ... in order, then a shuffle tree
// (a 30-min stream has 7 000 chunks: one thread walking them was 180 us).  stats[2b] = mean, [2b+1] = population variance
__global__ void wave_merge_stats_kernel(int n_utts, const int32_t *chunk_first, const double *partial, double *stats)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n_utts) return;
    const int c0 = chunk_first[b], c1 = chunk_first[b + 1];
    // contiguous runs per lane keep the merge order the order of the samples
    const int per = (c1 - c0 + 31) / 32;
    Moments m{0.0, 0.0, 0.0};
    for (int c = c0 + lane * per; c < c1 && c < c0 + (lane + 1) * per; ++c)
        m = merge(m, Moments{partial[3 * (size_t)c], partial[3 * (size_t)c + 1], partial[3 * (size_t)c + 2]});
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Moments o = shfl_xor(m, d);
        m = (lane & d) ? merge(o, m) : merge(m, o);
    }
    if (lane == 0) {
        stats[2 * b] = m.mean;
        stats[2 * b + 1] = (m.n > 0.0) ? m.m2 / m.n : 0.0;
    }
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(kStatThreads, sizeof(InT) + sizeof(OutT) == 8 ? 8 : 5)
wave_apply_norm_kernel(const InT *wave, OutT *out, const int64_t *n_samples, const int64_t *wave_off,
                       const int32_t *chunk_utt, const int32_t *chunk_first, const double *stats, int mode)
{
    const int utt = chunk_utt[blockIdx.x];
    const int64_t c = blockIdx.x - chunk_first[utt];
    const int64_t n = n_samples[utt];
    const int64_t j0 = c * kStatChunk;
    const int64_t len = (n - j0 < kStatChunk) ? n - j0 : kStatChunk;
    const InT *src = wave + wave_off[utt] + j0;
    OutT *dst = out + wave_off[utt] + j0;
    InT v[kStatPer];
    load16(src, len, threadIdx.x, v);
    alignas(16) OutT r[kStatPer];
    if (mode == 0) { // z-score, float64 arithmetic as numpy does for float64 input (correctly rounded quotient)
        const Znorm zn = Znorm::from_stats(stats, utt);
#pragma unroll
        for (int k = 0; k < kStatPer; ++k) r[k] = (OutT)zn((double)v[k]);
    } else { // wav2vec2 feature extractor: float32 arithmetic on the float32-rounded statistics
        const float mf = (float)stats[2 * utt];
        const float denom = sqrtf(__fadd_rn((float)stats[2 * utt + 1], 1e-7f));
#pragma unroll
        for (int k = 0; k < kStatPer; ++k) r[k] = (OutT)__fdiv_rn(__fsub_rn((float)v[k], mf), denom);
    }
    // r[] is in the INPUT type's slot order; when both types have the same width the 16-byte vectors line up and whole
    // chunks are stored as vectors, otherwise every value goes to its own position (still 4- or 8-byte coalesced)
    if (sizeof(InT) == sizeof(OutT) && len == kStatChunk && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        constexpr int kVec = 16 / (int)sizeof(OutT);
#pragma unroll
        for (int q = 0; q < kStatPer / kVec; ++q)
            reinterpret_cast<uint4 *>(dst)[q * kStatThreads + threadIdx.x] = reinterpret_cast<const uint4 *>(r)[q];
    } else {
#pragma unroll
        for (int s = 0; s < kStatPer; ++s) {
            const int i = slot_index<InT>(s, threadIdx.x);
            if (i < len) dst[i] = r[s];
        }
    }
}

// Same affine map, written straight into the feature extractor's padded layout: out[b, :n_b] = normalised samples,
// out[b, n_b:] = 0 (padding_value), mask[b, :] = 1 / 0 (ref:src/aat/training/collate.py:301-304: the processor call
// with padding=True).  grid = (chunks of the padded row, utterances).
template <typename InT>
__global__ void __launch_bounds__(kStatThreads)
wave_apply_norm_padded_kernel(const InT *__restrict__ wave, float *__restrict__ out, int32_t *__restrict__ mask,
                              const int64_t *n_samples, const int64_t *wave_off, int64_t n_max, const double *stats, int mode)
{
    const int utt = blockIdx.y;
    const int64_t n = n_samples[utt];
    const int64_t j0 = (int64_t)blockIdx.x * kStatChunk;
    const int64_t end = (n_max - j0 < kStatChunk) ? n_max : j0 + kStatChunk;
    const InT *src = wave + wave_off[utt];
    float *dst = out + (size_t)utt * n_max;
    int32_t *m = mask ? mask + (size_t)utt * n_max : nullptr;
    const Znorm zn = Znorm::from_stats(stats, utt);
    const float mf = (float)stats[2 * utt];
    const float denom32 = sqrtf(__fadd_rn((float)stats[2 * utt + 1], 1e-7f));
    auto norm = [&](int64_t i) -> float {
        if (i >= n) return 0.0f; // padding_value
        return (mode == 0) ? (float)zn((double)src[i]) : __fdiv_rn(__fsub_rn((float)src[i], mf), denom32);
    };
    // rows of the padded layout start 16-byte aligned when n_max is a multiple of 4: four samples per thread and trip,
    // one 16-byte store for the values and one for the mask (the packed source is read with coalesced scalar loads)
    const bool vec = (n_max & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                     (m == nullptr || (reinterpret_cast<uintptr_t>(mask) & 15) == 0);
    if (vec) {
        for (int64_t i = j0 + 4 * (int64_t)threadIdx.x; i < end; i += 4 * kStatThreads) {
            *reinterpret_cast<float4 *>(dst + i) = make_float4(norm(i), norm(i + 1), norm(i + 2), norm(i + 3));
            if (m) *reinterpret_cast<int4 *>(m + i) = make_int4(i < n, i + 1 < n, i + 2 < n, i + 3 < n);
        }
    } else {
        for (int64_t i = j0 + threadIdx.x; i < end; i += kStatThreads) {
            dst[i] = norm(i);
            if (m) m[i] = i < n ? 1 : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------- N1
__global__ void pad_boarders_kernel(int n_utts, int64_t s_max, const int64_t *seg_slot_off, const int64_t *seg_len,
                                    const int32_t *seg_count, int64_t *boarders, int64_t *mask, int32_t *status)
{
    // one warp per utterance: inclusive scan of the lengths (ref:src/aat/training/collate.py:158 `.cumsum()`)
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n_utts) return;
    const int cnt = seg_count[b];
    const int64_t *len = seg_len + seg_slot_off[b];
    int64_t *brow = boarders + (size_t)b * s_max, *mrow = mask + (size_t)b * s_max;
    if (lane == 0 && cnt > s_max) status[b] = AAT_ERR_CAPACITY;
    int64_t base = 0;
    for (int64_t i0 = 0; i0 < s_max; i0 += 32) {
        const int64_t i = i0 + lane;
        const int64_t v = (i < cnt) ? len[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (i < s_max) {
            brow[i] = (i < cnt) ? base + incl : 0;
            mrow[i] = (i < cnt) ? 1 : 0;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// grid = B * s_max rows.  Row (b, s): out[b, s, :len] = wave[b, prev:boarder], rest 0; mask likewise.
__global__ void __launch_bounds__(256)
scatter_segments_kernel(const float *wave, int64_t n_max, const int64_t *boarders, int64_t s_max, int64_t max_frames,
                        float *out, float *mask, int32_t *status)
{
    const int64_t row = blockIdx.x;
    const int64_t b = row / s_max, s = row - b * s_max;
    const int64_t *brow = boarders + b * s_max;
    const int64_t end = brow[s];
    float *o = out + (size_t)row * max_frames;
    float *m = mask ? mask + (size_t)row * max_frames : nullptr;
    int64_t len = 0, begin = 0;
    if (s == 0 || end != 0) { // `if segment_i > 0 and segment_boarder == 0: continue` (collate.py:326-327)
        begin = (s == 0) ? 0 : brow[s - 1];
        len = end - begin;
        // the reference asserts prev < boarder and fails on a shape mismatch when the slice is longer than the
        // tile or runs past the padded waveform; report instead of raising
        if (len <= 0 || len > max_frames || end > n_max) {
            if (threadIdx.x == 0) atomicMin(status + b, (int32_t)AAT_ERR_INVALID);
            if (len < 0) len = 0;
            if (len > max_frames) len = max_frames;
            if (begin + len > n_max) len = (n_max > begin) ? n_max - begin : 0;
        }
    }
    const float *src = wave + (size_t)b * n_max + begin;
    if ((max_frames & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
        (m == nullptr || (reinterpret_cast<uintptr_t>(mask) & 15) == 0)) {
        // rows are 16-byte aligned: 128-bit stores (the ragged source is read with scalar, coalesced loads)
        for (int64_t i = 4 * (int64_t)threadIdx.x; i < max_frames; i += 4 * (int64_t)blockDim.x) {
            float4 v, k;
            v.x = (i + 0 < len) ? src[i + 0] : 0.0f, k.x = (i + 0 < len) ? 1.0f : 0.0f;
            v.y = (i + 1 < len) ? src[i + 1] : 0.0f, k.y = (i + 1 < len) ? 1.0f : 0.0f;
            v.z = (i + 2 < len) ? src[i + 2] : 0.0f, k.z = (i + 2 < len) ? 1.0f : 0.0f;
            v.w = (i + 3 < len) ? src[i + 3] : 0.0f, k.w = (i + 3 < len) ? 1.0f : 0.0f;
            *reinterpret_cast<float4 *>(o + i) = v;
            if (m) *reinterpret_cast<float4 *>(m + i) = k;
        }
    } else {
        for (int64_t i = threadIdx.x; i < max_frames; i += blockDim.x) {
            const bool in = i < len;
            o[i] = in ? src[i] : 0.0f;
            if (m) m[i] = in ? 1.0f : 0.0f;
        }
    }
}

// grid = B * s_max rows; out[b, s, mel, :cols] = mel_b[mel, prev/hop : boarder/hop], rest 0
// Utterance b's mel is a C-contiguous (n_mels, T_b) block: either the plan's packed layout (frame_off / n_samples) or,
// when mel_elem_off / mel_frames are given, any blocks the caller describes (cropped mels of the n-word path).
__global__ void __launch_bounds__(256)
scatter_mel_segments_kernel(const float *__restrict__ mel, const int64_t *frame_off, const int64_t *n_samples,
                            const int64_t *mel_elem_off, const int64_t *mel_frames, const int64_t *mel_row_stride, int hop, int n_mels,
                            const int64_t *boarders, int64_t s_max, int64_t max_items, float *__restrict__ out, int32_t *status)
{
    // The tile's geometry (four 64-bit divisions, a chain of dependent table loads) is worked out by ONE thread: done
    // by all 256 it was two thirds of the kernel's instructions (ncu: 21 M warp instructions for 128 MB).
    __shared__ int64_t s_geo[4]; // source offset, row stride, columns to copy
    const int64_t row = blockIdx.x;
    if (threadIdx.x == 0) {
        const int64_t b = row / s_max, s = row - b * s_max;
        const int64_t *brow = boarders + b * s_max;
        const int64_t end = brow[s];
        const int64_t T = mel_frames ? mel_frames[b] : 1 + n_samples[b] / hop; // columns of the block (slices clamp here)
        int64_t c0 = 0, cols = 0;
        if (s == 0 || end != 0) {
            const int64_t begin = (s == 0) ? 0 : brow[s - 1];
            c0 = begin / hop;
            int64_t c1 = end / hop;
            if (c1 > T) c1 = T; // numpy slicing clamps at the array end
            if (c0 > T) c0 = T;
            cols = c1 - c0;
            if (cols > max_items) { // the reference fails on the shape mismatch
                atomicMin(status + b, (int32_t)AAT_ERR_INVALID);
                cols = max_items;
            }
            if (cols < 0) cols = 0;
        }
        s_geo[0] = (mel_elem_off ? mel_elem_off[b] : (int64_t)n_mels * frame_off[b]) + c0;
        s_geo[1] = mel_row_stride ? mel_row_stride[b] : T; // elements between the block's rows
        s_geo[2] = cols;
    }
    __syncthreads();
    const float *src = mel + s_geo[0];
    const int64_t stride = s_geo[1], cols = s_geo[2];
    float *o = out + (size_t)row * n_mels * max_items;
    // The tile [n_mels, max_items] is one contiguous, 16-byte aligned range: every thread writes 16-byte vectors of
    // consecutive tile elements (a row of 151 floats is not a multiple of 16 bytes, so a warp per row left every row
    // with ragged sectors at both ends: 0.49 of the HBM peak, profiles/r1_collate_bw.txt); the source elements of a
    // vector are consecutive in the mel row except where the vector wraps to the next row.
    const int items = (int)max_items, ncols = (int)cols, total = n_mels * items;
    if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        if (ncols == 0) { // a padding tile (no segment in this slot): zeros, no loads
            for (int e0 = 4 * (int)threadIdx.x; e0 < total; e0 += 4 * (int)blockDim.x)
                *reinterpret_cast<float4 *>(o + e0) = make_float4(0.f, 0.f, 0.f, 0.f);
            return;
        }
        // The kernel was bound by instruction issue (79 % issue utilisation, 31 M warp instructions for 128 MB: a
        // division, four wrap tests and 64-bit offsets per 16-byte vector; ncu r2_prof_scatter_mel): the row of a vector
        // now comes from one multiply-high, and the 35 of 38 vectors per row that do not wrap take a straight path.
        const unsigned magic = 0xFFFFFFFFu / (unsigned)items + 1u; // e / items == umulhi(e, magic) for e, items < 2^16
        const bool small = items >= 2 && total < 65536 && (int64_t)n_mels * stride < (int64_t)INT32_MAX;
        const int step = 4 * (int)blockDim.x, istride = (int)stride;
        for (int e = 4 * (int)threadIdx.x; e < total; e += step) {
            int r = small ? (int)__umulhi((unsigned)e, magic) : e / items;
            int c = e - r * items;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (small && c + 4 <= items) {
                const float *p = src + (r * istride + c);
                const int left = ncols - c; // columns of the segment at or behind c
                if (left >= 4) {
                    v[0] = __ldg(p), v[1] = __ldg(p + 1), v[2] = __ldg(p + 2), v[3] = __ldg(p + 3);
                } else if (left > 0) {
                    v[0] = __ldg(p);
                    if (left > 1) v[1] = __ldg(p + 1);
                    if (left > 2) v[2] = __ldg(p + 2);
                }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (c < ncols) v[u] = __ldg(src + ((size_t)r * stride + c));
                    if (++c == items) c = 0, ++r;
                }
            }
            *reinterpret_cast<float4 *>(o + e) = make_float4(v[0], v[1], v[2], v[3]);
        }
    } else {
        for (int e = (int)threadIdx.x; e < total; e += (int)blockDim.x) {
            const int r = e / items, c = e - r * items;
            o[e] = (c < ncols) ? src[(size_t)r * stride + c] : 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------- N4
// Masked mean over the valid frames of every row of the padded layout [R, L, D] (R = batch * segments):
// the `SegmentProjectionEnum.mean` branch the reference leaves as NotImplementedError
// (ref:src/aslm/modeling_aslm.py:258-259), under the frame mask of encode_audio (:195-218).
// One CTA per row; a thread owns one 16-byte column slab; only frames whose mask is set are read, four
// loads in flight per thread.  Rows without a valid frame give zeros and row_mask = 0.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
    static constexpr int kCols = 4;
    __device__ static void add(const void *p, float (&a)[4])
    {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(p));
        a[0] += x.x, a[1] += x.y, a[2] += x.z, a[3] += x.w;
    }
};
template <>
struct Vec16<__half> {
    static constexpr int kCols = 8;
    __device__ static void add(const void *p, float (&a)[8])
    {
        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(p));
        const __half2 *h = reinterpret_cast<const __half2 *>(&x);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            a[2 * i] += f.x, a[2 * i + 1] += f.y;
        }
    }
};
template <>
struct Vec16<__nv_bfloat16> {
    static constexpr int kCols = 8;
    __device__ static void add(const void *p, float (&a)[8])
    {
        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[2 * i] += __uint_as_float(w[i] << 16);
            a[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
masked_mean_pool_kernel(const unsigned char *emb, const int64_t *mask, int64_t seq_len, int dim, int slabs_per_row,
                        float *out, int64_t *row_mask)
{
    using V = Vec16<T>;
    constexpr int kCols = V::kCols;
    extern __shared__ int s_valid[]; // indices of the valid frames of this row
    __shared__ int s_count;
    const int64_t r = blockIdx.x;
    const int64_t *mrow = mask + r * seq_len;
    if (threadIdx.x == 0) { // seq_len is short (<= 74 frames at the reference's settings): a serial compaction is fine
        int c = 0;
        for (int64_t t = 0; t < seq_len; ++t)
            if (mrow[t] != 0) s_valid[c++] = (int)t;
        s_count = c;
        if (row_mask) row_mask[r] = c > 0;
    }
    __syncthreads();
    const int count = s_count;
    const size_t row_bytes = (size_t)dim * sizeof(T);
    const unsigned char *base = emb + (size_t)r * seq_len * row_bytes;
    for (int slab = threadIdx.x; slab < slabs_per_row; slab += blockDim.x) {
        float acc[4][kCols];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < kCols; ++k) acc[u][k] = 0.0f;
        const unsigned char *col = base + (size_t)slab * 16;
        int i = 0;
        for (; i + 4 <= count; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) V::add(col + (size_t)s_valid[i + u] * row_bytes, acc[u]);
        }
        for (; i < count; ++i) V::add(col + (size_t)s_valid[i] * row_bytes, acc[0]);
        float *o = out + (size_t)r * dim + (size_t)slab * kCols;
        const float n = (float)count;
#pragma unroll
        for (int k = 0; k < kCols; k += 4) {
            float4 v;
            v.x = count ? __fdiv_rn((acc[0][k] + acc[1][k]) + (acc[2][k] + acc[3][k]), n) : 0.0f;
            v.y = count ? __fdiv_rn((acc[0][k + 1] + acc[1][k + 1]) + (acc[2][k + 1] + acc[3][k + 1]), n) : 0.0f;
            v.z = count ? __fdiv_rn((acc[0][k + 2] + acc[1][k + 2]) + (acc[2][k + 2] + acc[3][k + 2]), n) : 0.0f;
            v.w = count ? __fdiv_rn((acc[0][k + 3] + acc[1][k + 3]) + (acc[2][k + 3] + acc[3][k + 3]), n) : 0.0f;
            *reinterpret_cast<float4 *>(o + k) = v;
        }
    }
}

} // namespace

int launch_masked_mean_pool(const void *emb, int emb_dtype, int64_t n_rows, int64_t seq_len, int32_t dim,
                            const int64_t *mask, float *out, int64_t *row_mask, cudaStream_t stream)
{
    int esize;
    switch (emb_dtype) {
    case AAT_F32: esize = 4; break;
    case AAT_F16:
    case AAT_BF16: esize = 2; break;
    default: AAT_REQUIRE(false, AAT_ERR_UNSUPPORTED, "aat_masked_mean_pool: embedding dtype must be F32, F16 or BF16");
    }
    const int64_t row_bytes = (int64_t)dim * esize;
    AAT_REQUIRE(row_bytes % 16 == 0, AAT_ERR_UNSUPPORTED, "aat_masked_mean_pool: dim * sizeof(element) must be a multiple of 16");
    AAT_REQUIRE((reinterpret_cast<uintptr_t>(emb) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, AAT_ERR_INVALID,
                "aat_masked_mean_pool: emb_dev and out_dev must be 16-byte aligned");
    AAT_REQUIRE(n_rows < (int64_t)INT32_MAX && seq_len < (1 << 20), AAT_ERR_UNSUPPORTED, "aat_masked_mean_pool: shape too large");
    if (n_rows == 0 || dim == 0) return AAT_OK;
    const int slabs = (int)(row_bytes / 16);
    int threads = ((slabs + 31) / 32) * 32;
    if (threads > 256) threads = 256;
    const size_t smem = sizeof(int) * (size_t)(seq_len ? seq_len : 1);
    const unsigned char *e = static_cast<const unsigned char *>(emb);
    if (emb_dtype == AAT_F32)
        masked_mean_pool_kernel<float><<<(unsigned)n_rows, threads, smem, stream>>>(e, mask, seq_len, dim, slabs, out, row_mask);
    else if (emb_dtype == AAT_F16)
        masked_mean_pool_kernel<__half><<<(unsigned)n_rows, threads, smem, stream>>>(e, mask, seq_len, dim, slabs, out, row_mask);
    else
        masked_mean_pool_kernel<__nv_bfloat16><<<(unsigned)n_rows, threads, smem, stream>>>(e, mask, seq_len, dim, slabs, out, row_mask);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

int launch_normalize(aat_ctx *ctx, const aat_plan *plan, const void *wave, int in_dtype, int mode, void *out,
                     int out_dtype, double *stats, cudaStream_t stream)
{
    (void)ctx;
    AAT_REQUIRE(in_dtype == AAT_F32 || in_dtype == AAT_F64, AAT_ERR_UNSUPPORTED, "aat_normalize: input dtype must be F32 or F64");
    AAT_REQUIRE(out_dtype == AAT_F32 || out_dtype == AAT_F64, AAT_ERR_UNSUPPORTED, "aat_normalize: output dtype must be F32 or F64");
    AAT_REQUIRE(mode == 0 || mode == 1, AAT_ERR_INVALID, "aat_normalize: unknown mode %d", mode);
    if (plan->norm_chunks == 0) return AAT_OK;
    double *st = stats ? stats : plan->d_norm_stats;
    if (in_dtype == AAT_F32)
        wave_chunk_stats_kernel<float><<<plan->norm_chunks, kStatThreads, 0, stream>>>(
            static_cast<const float *>(wave), plan->d_n_samples, plan->d_wave_off, plan->d_chunk_utt, plan->d_chunk_first,
            plan->d_norm_partial);
    else
        wave_chunk_stats_kernel<double><<<plan->norm_chunks, kStatThreads, 0, stream>>>(
            static_cast<const double *>(wave), plan->d_n_samples, plan->d_wave_off, plan->d_chunk_utt, plan->d_chunk_first,
            plan->d_norm_partial);
    AAT_LAUNCH_CHECK();
    wave_merge_stats_kernel<<<(plan->n_utts + 3) / 4, 128, 0, stream>>>(plan->n_utts, plan->d_chunk_first,
                                                                        plan->d_norm_partial, st);
    AAT_LAUNCH_CHECK();
    if (out == nullptr) return AAT_OK; // statistics only
#define AAT_APPLY(IN, OUT)                                                                                         \
    wave_apply_norm_kernel<IN, OUT><<<plan->norm_chunks, kStatThreads, 0, stream>>>(                                \
        static_cast<const IN *>(wave), static_cast<OUT *>(out), plan->d_n_samples, plan->d_wave_off, plan->d_chunk_utt, \
        plan->d_chunk_first, st, mode)
    if (in_dtype == AAT_F32 && out_dtype == AAT_F32) AAT_APPLY(float, float);
    else if (in_dtype == AAT_F32) AAT_APPLY(float, double);
    else if (out_dtype == AAT_F32) AAT_APPLY(double, float);
    else AAT_APPLY(double, double);
#undef AAT_APPLY
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

int launch_pad_boarders(const aat_plan *plan, const int64_t *seg_len, const int32_t *seg_count, int64_t s_max,
                        int64_t *boarders, int64_t *mask, int32_t *status, cudaStream_t stream)
{
    if (plan->n_utts == 0) return AAT_OK;
    AAT_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)plan->n_utts, stream));
    const int warps = 4; // s_max == 0 still runs: an utterance with segments must report AAT_ERR_CAPACITY
    pad_boarders_kernel<<<(plan->n_utts + warps - 1) / warps, warps * 32, 0, stream>>>(
        plan->n_utts, s_max, plan->d_seg_slot_off, seg_len, seg_count, boarders, mask, status);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

int launch_scatter_segments(const float *wave, int64_t n_max, int32_t n_utts, const int64_t *boarders, int64_t s_max,
                            int64_t max_frames, float *out, float *mask, int32_t *status, cudaStream_t stream)
{
    if (n_utts == 0) return AAT_OK;
    AAT_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)n_utts, stream)); // also on the empty shapes below
    if (s_max == 0 || max_frames == 0) return AAT_OK;
    AAT_REQUIRE((int64_t)n_utts * s_max < (int64_t)INT32_MAX, AAT_ERR_UNSUPPORTED, "aat_scatter_segments: too many rows");
    scatter_segments_kernel<<<(unsigned)(n_utts * s_max), 256, 0, stream>>>(wave, n_max, boarders, s_max, max_frames, out,
                                                                          mask, status);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

int launch_scatter_mel_segments(aat_ctx *ctx, const aat_plan *plan, int32_t n_utts, const float *mel,
                                const int64_t *mel_elem_off, const int64_t *mel_frames, const int64_t *mel_row_stride,
                                const int64_t *boarders,
                                int64_t s_max, int64_t max_items, float *out, int32_t *status, cudaStream_t stream)
{
    if (n_utts == 0) return AAT_OK;
    AAT_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)n_utts, stream)); // also on the empty shapes below
    if (s_max == 0 || max_items == 0) return AAT_OK;
    AAT_REQUIRE((int64_t)n_utts * s_max < (int64_t)INT32_MAX, AAT_ERR_UNSUPPORTED, "aat_scatter_mel_segments: too many rows");
    scatter_mel_segments_kernel<<<(unsigned)(n_utts * s_max), 256, 0, stream>>>(
        mel, plan ? plan->d_frame_off : nullptr, plan ? plan->d_n_samples : nullptr, mel_elem_off, mel_frames,
        mel_row_stride, ctx->cfg.hop_length, ctx->cfg.num_mel_filters, boarders, s_max, max_items, out, status);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

int launch_normalize_padded(aat_ctx *ctx, const aat_plan *plan, const void *wave, int in_dtype, int mode, float *out,
                            int64_t n_max, int32_t *mask, double *stats, cudaStream_t stream)
{
    AAT_REQUIRE(in_dtype == AAT_F32 || in_dtype == AAT_F64, AAT_ERR_UNSUPPORTED, "aat_normalize_padded: input dtype must be F32 or F64");
    AAT_REQUIRE(mode == 0 || mode == 1, AAT_ERR_INVALID, "aat_normalize_padded: unknown mode %d", mode);
    if (plan->n_utts == 0 || n_max == 0) return AAT_OK;
    double *st = stats ? stats : plan->d_norm_stats;
    const int rc = launch_normalize(ctx, plan, wave, in_dtype, mode, nullptr, AAT_F32, st, stream); // statistics only
    if (rc != AAT_OK) return rc;
    const dim3 grid((unsigned)((n_max + kStatChunk - 1) / kStatChunk), (unsigned)plan->n_utts);
    if (in_dtype == AAT_F32)
        wave_apply_norm_padded_kernel<float><<<grid, kStatThreads, 0, stream>>>(
            static_cast<const float *>(wave), out, mask, plan->d_n_samples, plan->d_wave_off, n_max, st, mode);
    else
        wave_apply_norm_padded_kernel<double><<<grid, kStatThreads, 0, stream>>>(
            static_cast<const double *>(wave), out, mask, plan->d_n_samples, plan->d_wave_off, n_max, st, mode);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat
