// K3: amplitude-minima boundary scan + merge/split state machine -> integer segment offsets.
//
// Replaces find_amplitude_minimas, pretokenize and process_segments_boarders
// (ref:src/aat/tokenizer.py:55-92, 121-139, 141-183).  The result is integer and must equal the
// reference's bit for bit, which pins the float32 arithmetic completely:
//   amp[t] = -10 * (sum_r mel[r][t], rows added in order, float32) / n_mels   numpy mean(axis=0)
//   cs[t]  = cs[t-1] + amp[t]                    sequential float32 recurrence numpy cumsum
//   rm[i]  = (cs[i+n] - cs[i]) / float(n)        float32
//   minimum at i  <=>  rm[i] > rm[i+1] + 1e-5f  and  rm[i] > rm[i-1] + 1e-5f  and  rm[i] > max_amp
//                      (argrelextrema order=1 mode='clip': the end points never qualify)
// Every operation uses an explicit round-to-nearest intrinsic so nothing is contracted into an FMA.
// A parallel prefix scan would change ~10 % of the minima on long audio (SURVEY.md §7 hard part 1),
// so the cumsum stays a serial chain on one thread; everything around it (column means, running mean,
// comparisons, ordered compaction) is thread-parallel.  One CTA per utterance, time axis in chunks.
#include "aat_internal.cuh"

namespace aat {

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096; // mel frames per pass; must be >= running_mean_points + 2

struct SegState {
    int64_t prev;
    int64_t count;
    int32_t status;   // < 0: aat_status error; otherwise bit 0 = the last segment is the zero-padded tail
    int64_t frames;   // HuBERT frames of the segments emitted so far (per-segment encode convention)
    int64_t *local;   // optional: local[i] = frames before segment i of this utterance
};

__device__ __forceinline__ int64_t hubert_frames(int64_t len)
{
    return (len < 400) ? 0 : (len - 400) / 320 + 1; // TF:models/hubert/modeling_hubert.py:675-688, closed form
}

__device__ __forceinline__ void emit_segment(int64_t start, int64_t len, int64_t *seg_start, int64_t *seg_len,
                                             int64_t capacity, SegState &st)
{
    if (st.count < capacity) {
        seg_start[st.count] = start;
        seg_len[st.count] = len;
        if (st.local) st.local[st.count] = st.frames;
        st.frames += hubert_frames(len);
    } else {
        st.status = AAT_ERR_CAPACITY;
    }
    ++st.count;
}

// One iteration of the loop at ref:src/aat/tokenizer.py:154-175.
__device__ __forceinline__ void push_boarder(int64_t b, int64_t min_frames, int64_t max_frames, int64_t *seg_start,
                                             int64_t *seg_len, int64_t capacity, SegState &st)
{
    const int64_t len = b - st.prev;
    if (len < min_frames) return; // merge forward: prev stays
    if (len > max_frames) {
        const int64_t k = len / max_frames;
        const int64_t gap = len - k * max_frames;
        int64_t n_cuts = k;
        int64_t last_cut = k * max_frames;
        if (gap == 0)
            n_cuts = k - 1; // drop last empty segment
        else if (gap < min_frames)
            last_cut = len - min_frames; // may fall below the previous cut when min > max
        int64_t lo = 0;
        for (int64_t j = 0; j < n_cuts; ++j) { // np.split: a[lo:c] with Python slice clamping
            const int64_t c = (j == k - 1) ? last_cut : (j + 1) * max_frames;
            const int64_t a0 = lo < len ? lo : len;
            const int64_t a1 = c < len ? c : len;
            emit_segment(st.prev + a0, a1 > a0 ? a1 - a0 : 0, seg_start, seg_len, capacity, st);
            lo = c;
        }
        const int64_t a0 = lo < len ? lo : len;
        emit_segment(st.prev + a0, len - a0, seg_start, seg_len, capacity, st);
    } else {
        emit_segment(st.prev, len, seg_start, seg_len, capacity, st);
    }
    st.prev = b;
}

// ref:src/aat/tokenizer.py:177-181: zero-padded tail of min_segment_frames samples.
__device__ __forceinline__ void finish_segments(int64_t n_samples, int64_t min_frames, int64_t *seg_start,
                                                int64_t *seg_len, int64_t capacity, SegState &st)
{
    if (st.prev != n_samples) {
        if (st.status == 0) st.status = (n_samples - st.prev > min_frames) ? AAT_ERR_TAIL : 1;
        emit_segment(st.prev, min_frames, seg_start, seg_len, capacity, st);
    }
}


__device__ __forceinline__ int64_t ld_cg(const int64_t *p) { return __ldcg(reinterpret_cast<const long long *>(p)); }

// Segment lengths -> packed CSR of HuBERT frame offsets (per-segment encode convention), by ONE CTA.
// s_seg / s_frm: shared scratch of n_utts + 1 int64 each.  Loads bypass L1 (the lengths may have been
// written by other CTAs of the same launch).
__device__ void build_frame_csr(int n_utts, const int64_t *seg_slot_off, const int64_t *seg_len,
                                const int32_t *seg_count, int64_t *seg_off, int64_t *n_seg_out,
                                int64_t *utt_seg_off_out, int64_t *s_seg, int64_t *s_frm)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_threads = blockDim.x, n_warps = n_threads >> 5;
    __shared__ int64_t s_wsum[2][32];
    // per-utterance totals (a warp per utterance)
    for (int b = warp; b < n_utts; b += n_warps) {
        const int64_t *len = seg_len + seg_slot_off[b];
        const int cnt = __ldcg(seg_count + b);
        int64_t sum = 0;
        for (int i = lane; i < cnt; i += 32) sum += hubert_frames(ld_cg(len + i));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (lane == 0) {
            s_seg[b + 1] = cnt;
            s_frm[b + 1] = sum;
        }
    }
    if (tid == 0) s_seg[0] = 0, s_frm[0] = 0;
    __syncthreads();
    // block-wide inclusive scan of both arrays: contiguous chunk per thread, shuffle scan of the chunk totals
    const int per = (n_utts + n_threads - 1) / n_threads;
    const int b0 = 1 + tid * per, b1 = min(n_utts + 1, b0 + per);
    int64_t a = 0, f = 0;
    for (int b = b0; b < b1; ++b) a += s_seg[b], f += s_frm[b];
    int64_t ia = a, jf = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t va = __shfl_up_sync(0xffffffffu, ia, d), vf = __shfl_up_sync(0xffffffffu, jf, d);
        if (lane >= d) ia += va, jf += vf;
    }
    if (lane == 31) s_wsum[0][warp] = ia, s_wsum[1][warp] = jf;
    __syncthreads();
    if (warp == 0) {
        int64_t wa = lane < n_warps ? s_wsum[0][lane] : 0, wf = lane < n_warps ? s_wsum[1][lane] : 0;
        int64_t xa = wa, xf = wf;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t va = __shfl_up_sync(0xffffffffu, xa, d), vf = __shfl_up_sync(0xffffffffu, xf, d);
            if (lane >= d) xa += va, xf += vf;
        }
        s_wsum[0][lane] = xa - wa, s_wsum[1][lane] = xf - wf; // exclusive warp offsets
    }
    __syncthreads();
    {
        int64_t ra = s_wsum[0][warp] + ia - a, rf = s_wsum[1][warp] + jf - f; // exclusive prefix of this chunk
        for (int b = b0; b < b1; ++b) {
            ra += s_seg[b], rf += s_frm[b];
            s_seg[b] = ra, s_frm[b] = rf;
        }
    }
    __syncthreads();
    if (tid == 0) {
        *n_seg_out = s_seg[n_utts];
        seg_off[s_seg[n_utts]] = s_frm[n_utts];
    }
    if (utt_seg_off_out)
        for (int b = tid; b <= n_utts; b += n_threads) utt_seg_off_out[b] = s_seg[b];
    // running offsets inside each utterance (a warp per utterance)
    for (int b = warp; b < n_utts; b += n_warps) {
        const int64_t *len = seg_len + seg_slot_off[b];
        const int cnt = __ldcg(seg_count + b);
        int64_t base = s_frm[b];
        int64_t *dst = seg_off + s_seg[b];
        for (int i0 = 0; i0 < cnt; i0 += 32) {
            const int i = i0 + lane;
            const int64_t fr = (i < cnt) ? hubert_frames(ld_cg(len + i)) : 0;
            int64_t incl = fr;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (i < cnt) dst[i] = base + incl - fr;
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// Fused form used by the boundaries kernel's last CTA: every utterance's CTA has already written its frame
// total and the local (within-utterance) frame offset of each segment, so what is left is one scan over the
// utterances and a rebase — two rounds of independent loads instead of a per-utterance dependency chain.
__device__ void rebase_frame_csr(int n_utts, const int64_t *seg_slot_off, const int64_t *seg_local,
                                 const int64_t *utt_frames, const int32_t *seg_count, int64_t *seg_off,
                                 int64_t *n_seg_out, int64_t *utt_seg_off_out, int64_t *s_seg, int64_t *s_frm)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_threads = blockDim.x, n_warps = n_threads >> 5;
    __shared__ int64_t s_wsum[2][32];
    for (int b = tid; b < n_utts; b += n_threads) {
        s_seg[b + 1] = __ldcg(seg_count + b);
        s_frm[b + 1] = ld_cg(utt_frames + b);
    }
    if (tid == 0) s_seg[0] = 0, s_frm[0] = 0;
    __syncthreads();
    const int per = (n_utts + n_threads - 1) / n_threads;
    const int b0 = 1 + tid * per, b1 = min(n_utts + 1, b0 + per);
    int64_t a = 0, f = 0;
    for (int b = b0; b < b1; ++b) a += s_seg[b], f += s_frm[b];
    int64_t ia = a, jf = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t va = __shfl_up_sync(0xffffffffu, ia, d), vf = __shfl_up_sync(0xffffffffu, jf, d);
        if (lane >= d) ia += va, jf += vf;
    }
    if (lane == 31) s_wsum[0][warp] = ia, s_wsum[1][warp] = jf;
    __syncthreads();
    if (warp == 0) {
        int64_t wa = lane < n_warps ? s_wsum[0][lane] : 0, wf = lane < n_warps ? s_wsum[1][lane] : 0;
        int64_t xa = wa, xf = wf;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t va = __shfl_up_sync(0xffffffffu, xa, d), vf = __shfl_up_sync(0xffffffffu, xf, d);
            if (lane >= d) xa += va, xf += vf;
        }
        s_wsum[0][lane] = xa - wa, s_wsum[1][lane] = xf - wf;
    }
    __syncthreads();
    {
        int64_t ra = s_wsum[0][warp] + ia - a, rf = s_wsum[1][warp] + jf - f;
        for (int b = b0; b < b1; ++b) {
            ra += s_seg[b], rf += s_frm[b];
            s_seg[b] = ra, s_frm[b] = rf;
        }
    }
    __syncthreads();
    if (tid == 0) {
        *n_seg_out = s_seg[n_utts];
        seg_off[s_seg[n_utts]] = s_frm[n_utts];
    }
    if (utt_seg_off_out)
        for (int b = tid; b <= n_utts; b += n_threads) utt_seg_off_out[b] = s_seg[b];
    // rebase: half-warps take utterances round-robin; loads of successive utterances are independent
    const int half = tid >> 4, hl = tid & 15, n_halves = n_threads >> 4;
#pragma unroll 4
    for (int b = half; b < n_utts; b += n_halves) {
        const int cnt = (int)(s_seg[b + 1] - s_seg[b]);
        const int64_t *src = seg_local + seg_slot_off[b];
        int64_t *dst = seg_off + s_seg[b];
        const int64_t base = s_frm[b];
        for (int i = hl; i < cnt; i += 16) dst[i] = base + ld_cg(src + i);
    }
}

struct BoundaryParams {
    const float *mel;
    const float *amp;
    const int64_t *n_samples;
    const int64_t *frame_off;
    const int64_t *seg_slot_off;
    int64_t *seg_start;
    int64_t *seg_len;
    int32_t *seg_count;
    int64_t *minima;
    int32_t *minima_count;
    int32_t *status;
    // optional fused frame CSR (built by the last CTA to finish)
    int64_t *seg_off;
    int64_t *n_seg;
    int64_t *utt_seg_off;
    int64_t *seg_local;   // [total_seg_slots] plan scratch: within-utterance frame offsets
    int64_t *utt_frames;  // [n_utts] plan scratch: frames per utterance
    unsigned *ticket;
    int n_utts;
    int64_t min_frames, max_frames;
    int hop, n_mels, npts;
    float max_amp;
};

__global__ void __launch_bounds__(kThreads) boundaries_kernel(const BoundaryParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int halo = p.npts + 2;
    const int halo_pad = (halo + 3) & ~3; // keeps the chunk part of s_cs 16-byte aligned
    float *s_amp = reinterpret_cast<float *>(smem_raw);     // [kChunk]
    float *s_cs = s_amp + kChunk;                            // [halo_pad + kChunk]; cs index g at [halo_pad + g - j0]
    int *s_min = reinterpret_cast<int *>(s_cs + halo_pad + kChunk); // [kChunk] minima of this pass (global frame index)
    __shared__ int s_warp_count[kThreads / 32];
    __shared__ int s_total;
    __shared__ float s_carry;

    const int tid = threadIdx.x;
    const int utt = blockIdx.x;
    const int64_t n = p.n_samples[utt];
    const int64_t T = 1 + n / p.hop;
    const int64_t fbase = p.frame_off[utt];
    const float *mel = p.mel ? p.mel + (size_t)p.n_mels * fbase : nullptr;
    const float *amp_in = p.amp ? p.amp + fbase : nullptr;
    const int64_t slot0 = p.seg_slot_off[utt];
    const int64_t capacity = p.seg_slot_off[utt + 1] - slot0;
    int64_t *seg_start = p.seg_start + slot0;
    int64_t *seg_len = p.seg_len + slot0;
    int64_t *minima_out = p.minima ? p.minima + fbase : nullptr;

    const int64_t L = T - p.npts; // running-mean length; minima live in [1, L-2]
    const float nf = (float)p.npts;
    SegState st{0, 0, 0, 0, p.seg_off ? p.seg_local + slot0 : nullptr};
    int64_t n_minima = 0;
    int64_t i_done = 1; // next candidate index to test

    for (int64_t j0 = 0; j0 < T; j0 += kChunk) {
        const int64_t j1 = (j0 + kChunk < T) ? j0 + kChunk : T;
        const int len = (int)(j1 - j0);

        // A: amplitude curve of this chunk (thread-parallel over time, sequential down the mel rows)
        for (int i = tid; i < len; i += kThreads) {
            const int64_t t = j0 + i;
            float a;
            if (amp_in) {
                a = amp_in[t];
            } else {
                float acc = mel[t];
#pragma unroll 8
                for (int r = 1; r < p.n_mels; ++r) acc = __fadd_rn(acc, mel[(size_t)r * T + t]);
                a = __fmul_rn(-10.0f, __fdiv_rn(acc, (float)p.n_mels));
            }
            s_amp[i] = a;
        }
        __syncthreads();

        // B: the serial float32 cumsum.  One thread; values are fetched 16 at a time into registers so the
        // shared-memory latency is paid once per batch and the chain itself runs at the FADD latency.
        if (tid == 0) {
            float run = (j0 == 0) ? 0.0f : s_carry;
            const float4 *src = reinterpret_cast<const float4 *>(s_amp);
            float4 *dst = reinterpret_cast<float4 *>(s_cs + halo_pad);
            int i = 0;
            for (; i + 16 <= len; i += 16) {
                float4 a = src[i / 4], b = src[i / 4 + 1], c4 = src[i / 4 + 2], d = src[i / 4 + 3];
                a.x = (j0 == 0 && i == 0) ? a.x : __fadd_rn(run, a.x);
                a.y = __fadd_rn(a.x, a.y), a.z = __fadd_rn(a.y, a.z), a.w = __fadd_rn(a.z, a.w);
                b.x = __fadd_rn(a.w, b.x), b.y = __fadd_rn(b.x, b.y), b.z = __fadd_rn(b.y, b.z), b.w = __fadd_rn(b.z, b.w);
                c4.x = __fadd_rn(b.w, c4.x), c4.y = __fadd_rn(c4.x, c4.y), c4.z = __fadd_rn(c4.y, c4.z), c4.w = __fadd_rn(c4.z, c4.w);
                d.x = __fadd_rn(c4.w, d.x), d.y = __fadd_rn(d.x, d.y), d.z = __fadd_rn(d.y, d.z), d.w = __fadd_rn(d.z, d.w);
                dst[i / 4] = a, dst[i / 4 + 1] = b, dst[i / 4 + 2] = c4, dst[i / 4 + 3] = d;
                run = d.w;
            }
            for (; i < len; ++i) {
                run = (j0 == 0 && i == 0) ? s_amp[0] : __fadd_rn(run, s_amp[i]);
                s_cs[halo_pad + i] = run;
            }
            s_carry = run;
        }
        __syncthreads();

        // C: running mean + strict-local-maximum test for every index whose neighbourhood is complete
        int64_t i_hi = j1 - p.npts - 1; // exclusive: needs cs[i + 1 + npts] < j1
        if (i_hi > L - 1) i_hi = L - 1;
        const int64_t range = i_hi > i_done ? i_hi - i_done : 0;
        const int per = (int)((range + kThreads - 1) / kThreads);
        unsigned mask = 0;
        {
            const int64_t a = i_done + (int64_t)tid * per;
            const float *cs = s_cs + halo_pad - j0;
            for (int u = 0; u < per; ++u) {
                const int64_t i = a + u;
                if (i >= i_hi) break;
                const float rl = __fdiv_rn(__fsub_rn(cs[i - 1 + p.npts], cs[i - 1]), nf);
                const float rc = __fdiv_rn(__fsub_rn(cs[i + p.npts], cs[i]), nf);
                const float rr = __fdiv_rn(__fsub_rn(cs[i + 1 + p.npts], cs[i + 1]), nf);
                const bool is_min = rc > __fadd_rn(rr, 1e-5f) && rc > __fadd_rn(rl, 1e-5f) && rc > p.max_amp;
                if (is_min) mask |= 1u << u;
            }
        }
        // ordered compaction: exclusive scan of per-thread counts
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += v;
        }
        if ((tid & 31) == 31) s_warp_count[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            int w = (tid < kThreads / 32) ? s_warp_count[tid] : 0;
            int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, wi, d);
                if (tid >= d) wi += v;
            }
            if (tid < kThreads / 32) s_warp_count[tid] = wi - w;
            if (tid == kThreads / 32 - 1) s_total = wi;
        }
        __syncthreads();
        {
            int pos = s_warp_count[tid >> 5] + incl - cnt;
            const int64_t a = i_done + (int64_t)tid * per;
            unsigned m = mask;
            while (m) {
                const int u = __ffs(m) - 1;
                m &= m - 1;
                s_min[pos++] = (int)(a + u);
            }
        }
        __syncthreads();
        const int found = s_total;

        // D: publish minima; one thread advances the merge/split state machine over the new boarders
        if (minima_out)
            for (int i = tid; i < found; i += kThreads) minima_out[n_minima + i] = s_min[i];
        if (tid == 0)
            for (int i = 0; i < found; ++i)
                push_boarder((int64_t)s_min[i] * p.hop, p.min_frames, p.max_frames, seg_start, seg_len, capacity, st);
        n_minima += found;
        if (i_hi > i_done) i_done = i_hi;

        // carry the last `halo` cumsum values into the next pass
        // (source [kChunk, kChunk + halo) and destination [0, halo) are disjoint because kChunk >= halo)
        if (j1 < T) {
            for (int i = tid; i < halo_pad; i += kThreads) s_cs[i] = s_cs[kChunk + i];
            __syncthreads();
        }
    }

    if (tid == 0) {
        push_boarder(n, p.min_frames, p.max_frames, seg_start, seg_len, capacity, st); // ref:src/aat/tokenizer.py:137
        finish_segments(n, p.min_frames, seg_start, seg_len, capacity, st);
        p.seg_count[utt] = (int32_t)(st.count < capacity ? st.count : capacity);
        if (p.minima_count) p.minima_count[utt] = (int32_t)n_minima;
        p.status[utt] = st.status;
        if (p.seg_off) p.utt_frames[utt] = st.frames;
    }

    // ---- fused epilogue: the last CTA to finish turns all segment lengths into the packed frame CSR ----
    if (p.seg_off != nullptr) {
        __shared__ int s_last;
        if (tid == 0) {
            __threadfence(); // this CTA's segment writes (all by thread 0) are visible before the ticket
            s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            int64_t *s_seg = reinterpret_cast<int64_t *>(smem_raw); // the chunk buffers are free now
            int64_t *s_frm = s_seg + p.n_utts + 1;
            rebase_frame_csr(p.n_utts, p.seg_slot_off, p.seg_local, p.utt_frames, p.seg_count, p.seg_off, p.n_seg,
                             p.utt_seg_off, s_seg, s_frm);
            if (tid == 0) *p.ticket = 0; // ready for the next launch (graph replays included)
        }
    }
}

__global__ void process_boarders_kernel(int64_t n_samples, const int64_t *boarders, int64_t n_boarders,
                                        int64_t min_frames, int64_t max_frames, int64_t *seg_start, int64_t *seg_len,
                                        int64_t capacity, int32_t *seg_count, int32_t *status)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    SegState st{0, 0, 0, 0, nullptr};
    for (int64_t i = 0; i < n_boarders; ++i)
        push_boarder(boarders[i], min_frames, max_frames, seg_start, seg_len, capacity, st);
    finish_segments(n_samples, min_frames, seg_start, seg_len, capacity, st);
    *seg_count = (int32_t)(st.count < capacity ? st.count : capacity);
    *status = st.status;
}

// Stand-alone form of the same CSR construction (one CTA).
constexpr int kCsrThreads = 1024;

__global__ void __launch_bounds__(kCsrThreads)
segment_frame_csr_kernel(int n_utts, const int64_t *seg_slot_off, const int64_t *seg_len, const int32_t *seg_count,
                         int64_t *seg_off, int64_t *n_seg_out, int64_t *utt_seg_off_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int64_t *s_seg = reinterpret_cast<int64_t *>(smem_raw);
    build_frame_csr(n_utts, seg_slot_off, seg_len, seg_count, seg_off, n_seg_out, utt_seg_off_out, s_seg,
                    s_seg + n_utts + 1);
}

} // namespace

int launch_boundaries(aat_ctx *ctx, const aat_plan *plan, const float *mel, const float *amp, int64_t *seg_start,
                      int64_t *seg_len, int32_t *seg_count, int64_t *minima, int32_t *minima_count, int32_t *status,
                      int64_t *seg_off, int64_t *n_seg, int64_t *utt_seg_off, cudaStream_t stream)
{
    if (plan->n_utts == 0) return AAT_OK;
    BoundaryParams p{};
    p.mel = mel;
    p.amp = amp;
    p.n_samples = plan->d_n_samples;
    p.frame_off = plan->d_frame_off;
    p.seg_slot_off = plan->d_seg_slot_off;
    p.seg_start = seg_start;
    p.seg_len = seg_len;
    p.seg_count = seg_count;
    p.minima = minima;
    p.minima_count = minima_count;
    p.status = status;
    p.min_frames = ctx->cfg.min_segment_frames;
    p.max_frames = ctx->cfg.max_segment_frames;
    p.hop = ctx->cfg.hop_length;
    p.n_mels = ctx->cfg.num_mel_filters;
    p.npts = ctx->cfg.running_mean_points;
    p.max_amp = ctx->cfg.max_amplitude_for_minima;
    p.n_utts = plan->n_utts;
    p.ticket = ctx->ticket;
    const size_t csr_smem = sizeof(int64_t) * 2 * (size_t)(plan->n_utts + 1);
    const bool fuse_csr = seg_off != nullptr && csr_smem <= sizeof(float) * 2 * kChunk;
    p.seg_off = fuse_csr ? seg_off : nullptr;
    p.n_seg = n_seg;
    p.utt_seg_off = utt_seg_off;
    p.seg_local = plan->d_seg_local;
    p.utt_frames = plan->d_utt_frames;
    const size_t smem = sizeof(float) * (size_t)(kChunk + ((p.npts + 2 + 3) & ~3) + kChunk) + sizeof(int) * (size_t)kChunk;
    AAT_CUDA_CHECK(cudaFuncSetAttribute(boundaries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ProfileScope prof(ctx, AAT_K_BOUNDARIES, stream);
        boundaries_kernel<<<plan->n_utts, kThreads, smem, stream>>>(p);
        AAT_LAUNCH_CHECK();
    }
    if (seg_off != nullptr && !fuse_csr) // batch too large for the fused epilogue's scratch: separate kernel
        return launch_segment_frame_csr(ctx, plan, seg_len, seg_count, seg_off, n_seg, utt_seg_off, stream);
    return AAT_OK;
}

int launch_process_boarders(aat_ctx *ctx, int64_t n_samples, const int64_t *boarders, int64_t n_boarders,
                            int64_t *seg_start, int64_t *seg_len, int64_t capacity, int32_t *seg_count,
                            int32_t *status, cudaStream_t stream)
{
    process_boarders_kernel<<<1, 32, 0, stream>>>(n_samples, boarders, n_boarders, ctx->cfg.min_segment_frames,
                                                  ctx->cfg.max_segment_frames, seg_start, seg_len, capacity,
                                                  seg_count, status);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

int launch_segment_frame_csr(aat_ctx *ctx, const aat_plan *plan, const int64_t *seg_len, const int32_t *seg_count,
                             int64_t *seg_off, int64_t *n_seg, int64_t *utt_seg_off, cudaStream_t stream)
{
    const size_t smem = sizeof(int64_t) * 2 * (size_t)(plan->n_utts + 1);
    AAT_REQUIRE(smem <= 200 * 1024, AAT_ERR_UNSUPPORTED, "aat_segment_frame_csr: at most 12799 utterances per plan");
    AAT_CUDA_CHECK(
        cudaFuncSetAttribute(segment_frame_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfileScope prof(ctx, AAT_K_FRAME_CSR, stream);
    segment_frame_csr_kernel<<<1, kCsrThreads, smem, stream>>>(plan->n_utts, plan->d_seg_slot_off, seg_len, seg_count,
                                                               seg_off, n_seg, utt_seg_off);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat
