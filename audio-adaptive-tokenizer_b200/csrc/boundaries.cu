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
// comparisons, ordered compaction) is thread-parallel and runs concurrently with the chain in a four-stage
// software pipeline.  One CTA per utterance.
#include "aat_internal.cuh"

namespace aat {

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 512; // mel frames per pipeline stage
// Tried and rejected (gpurun r2_b4): capping the kernel at 96 registers so that a CTA fits beside two log-mel CTAs
// (for the pipelined schedule of aat_b200/pipeline.py) slowed the serial scan of 30-min streams from 628 to 678 us and
// bought no overlap — where the log-mel kernel's free slots end up is not under our control.

// The merge/split state machine runs on ONE thread and is sequential in `prev`, so its cost per boarder is
// instruction latency.  It is templated on the index type: 32-bit arithmetic (no carry chains, single-instruction
// compares, multiply-high divisions) whenever every quantity fits in 31 bits — any utterance shorter than
// 37 hours at 16 kHz — and 64-bit otherwise.  Outputs are always written as int64.
template <typename I>
struct SegStateT {
    I prev;
    I count;
    int32_t status;   // < 0: aat_status error; otherwise bit 0 = the last segment is the zero-padded tail
    I frames;         // HuBERT frames of the segments emitted so far (per-segment encode convention)
    int64_t *local;   // optional: local[i] = frames before segment i of this utterance
};
using SegState = SegStateT<int64_t>;

template <typename I>
__device__ __forceinline__ I hubert_frames_t(I len)
{
    // TF:models/hubert/modeling_hubert.py:675-688, closed form
    return (len < 400) ? (I)0 : (I)((len - 400) / 320 + 1);
}
__device__ __forceinline__ int64_t hubert_frames(int64_t len)
{
    if (len <= 0x7fffffffLL) return (int64_t)hubert_frames_t<int32_t>((int32_t)len);
    return hubert_frames_t<int64_t>(len);
}

template <typename I>
__device__ __forceinline__ void emit_segment(I start, I len, int64_t *seg_start, int64_t *seg_len, I capacity,
                                             SegStateT<I> &st)
{
    if (st.count < capacity) {
        seg_start[st.count] = (int64_t)start;
        seg_len[st.count] = (int64_t)len;
        if (st.local) st.local[st.count] = (int64_t)st.frames;
        st.frames += hubert_frames_t<I>(len);
    } else {
        st.status = AAT_ERR_CAPACITY;
    }
    ++st.count;
}

// One iteration of the loop at ref:src/aat/tokenizer.py:154-175.
template <typename I>
__device__ __forceinline__ void push_boarder(I b, I min_frames, I max_frames, int64_t *seg_start, int64_t *seg_len,
                                             I capacity, SegStateT<I> &st)
{
    const I len = b - st.prev;
    if (len < min_frames) return; // merge forward: prev stays
    if (len > max_frames) {
        const I k = len / max_frames;
        const I gap = len - k * max_frames;
        I n_cuts = k;
        I last_cut = k * max_frames;
        if (gap == 0)
            n_cuts = k - 1; // drop last empty segment
        else if (gap < min_frames)
            last_cut = len - min_frames; // may fall below the previous cut when min > max
        I lo = 0;
        for (I j = 0; j < n_cuts; ++j) { // np.split: a[lo:c] with Python slice clamping
            const I c = (j == k - 1) ? last_cut : (I)((j + 1) * max_frames);
            const I a0 = lo < len ? lo : len;
            const I a1 = c < len ? c : len;
            emit_segment<I>(st.prev + a0, a1 > a0 ? (I)(a1 - a0) : (I)0, seg_start, seg_len, capacity, st);
            lo = c;
        }
        const I a0 = lo < len ? lo : len;
        emit_segment<I>(st.prev + a0, len - a0, seg_start, seg_len, capacity, st);
    } else {
        emit_segment<I>(st.prev, len, seg_start, seg_len, capacity, st);
    }
    st.prev = b;
}

// ref:src/aat/tokenizer.py:177-181: zero-padded tail of min_segment_frames samples.
template <typename I>
__device__ __forceinline__ void finish_segments(I n_samples, I min_frames, int64_t *seg_start, int64_t *seg_len,
                                                I capacity, SegStateT<I> &st)
{
    if (st.prev != n_samples) {
        if (st.status == 0) st.status = (n_samples - st.prev > min_frames) ? AAT_ERR_TAIL : 1;
        emit_segment<I>(st.prev, min_frames, seg_start, seg_len, capacity, st);
    }
}


constexpr unsigned long long kTotalsReady = 1ull << 63;
// polling load: relaxed but a real memory operation every time (never hoisted, never served from L1)
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
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
        n_seg_out[0] = s_seg[n_utts];
        n_seg_out[1] = s_frm[n_utts]; // the frames the CSR covers (AAT_POOL_ROWS_FROM_DEVICE)
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


// ---------------------------------------------------------------------------------------------------------------
// Two-phase form of the merge/split loop used by the boundary kernel.  The decisions are sequential in `prev`
// (ref:src/aat/tokenizer.py:154-175), but writing a segment — three 64-bit stores with their address arithmetic, the
// HuBERT frame count, the running frame offset — is not: done by the deciding thread it cost ~40 dependent
// instructions per boarder (2.1 us per 512-frame chunk, the bottleneck stage of the pipeline,
// profiles/r1_bnd_timeline_lookback.txt).  Now lane 0 of the emit warp only appends (start, length) pairs to a
// shared-memory queue, and the whole warp writes the queue out: coalesced stores, frame offsets by a warp scan.
constexpr int kSegQueue = 128; // (start, length) pairs per flush

template <typename I>
struct MergeSplit {
    I prev = 0;       // last accepted boarder
    // np.split of one over-long gap, resumable across queue flushes (pieces j = 0 .. n_cuts)
    bool splitting = false;
    I len = 0, k = 0, last_cut = 0, lo = 0, j = 0, n_cuts = 0, pending = 0;
};

// Appends segments to q[n ...] until the queue is full or the boarders [i, found) are used up; returns the new n.
// next(i) yields boarder i in samples.
template <typename I, typename Next>
__device__ __forceinline__ int fill_segment_queue(MergeSplit<I> &ms, int &i, const int found, Next next, const I min_frames,
                                                  const I max_frames, longlong2 *q)
{
    int n = 0;
    // fast path: boarders that are merged or accepted whole — no data-dependent branch (a lone warp pays ~20 cycles
    // for every taken one): the pair is stored unconditionally and the queue length advances only when it is accepted
    if (!ms.splitting) {
        I prev = ms.prev;
        bool long_gap = false;
        auto step = [&](const I b) { // false: over-long gap, the resumable loop below cuts it up
            const I len = b - prev;
            if (len > max_frames) return false;
            const bool accept = len >= min_frames;
            q[n] = make_longlong2((long long)prev, (long long)len);
            n += accept ? 1 : 0;
            prev = accept ? b : prev;
            ++i;
            return true;
        };
        // four boarders per trip: their loads are independent of the decisions and issue together
        while (i + 4 <= found && n + 4 <= kSegQueue) {
            const I b0 = next(i), b1 = next(i + 1), b2 = next(i + 2), b3 = next(i + 3);
            if (!(step(b0) && step(b1) && step(b2) && step(b3))) {
                long_gap = true;
                break;
            }
        }
        while (!long_gap && i < found && n < kSegQueue)
            if (!step(next(i))) break;
        ms.prev = prev;
    }
    while (n < kSegQueue) {
        if (ms.splitting) { // np.split: a[lo:c] with Python slice clamping, then the remainder
            if (ms.j < ms.n_cuts) {
                const I c = (ms.j == ms.k - 1) ? ms.last_cut : (I)((ms.j + 1) * max_frames);
                const I a0 = ms.lo < ms.len ? ms.lo : ms.len;
                const I a1 = c < ms.len ? c : ms.len;
                q[n++] = make_longlong2((long long)(ms.prev + a0), (long long)(a1 > a0 ? (I)(a1 - a0) : (I)0));
                ms.lo = c;
                ++ms.j;
            } else {
                const I a0 = ms.lo < ms.len ? ms.lo : ms.len;
                q[n++] = make_longlong2((long long)(ms.prev + a0), (long long)(ms.len - a0));
                ms.prev = ms.pending;
                ms.splitting = false;
            }
            continue;
        }
        if (i == found) break;
        const I b = next(i++);
        const I len = b - ms.prev;
        if (len < min_frames) continue; // merge forward: prev stays
        if (len > max_frames) {
            ms.len = len;
            ms.k = len / max_frames;
            const I gap = len - ms.k * max_frames;
            ms.n_cuts = ms.k;
            ms.last_cut = ms.k * max_frames;
            if (gap == 0)
                ms.n_cuts = ms.k - 1; // drop last empty segment
            else if (gap < min_frames)
                ms.last_cut = len - min_frames; // may fall below the previous cut when min > max
            ms.lo = 0, ms.j = 0, ms.pending = b;
            ms.splitting = true;
            continue;
        }
        q[n++] = make_longlong2((long long)ms.prev, (long long)len);
        ms.prev = b;
    }
    return n;
}

// The whole warp writes q[0, n) behind the `count` segments already out; count / frames / status are warp-uniform.
__device__ __forceinline__ void flush_segment_queue(const longlong2 *q, const int n, const int lane, int64_t *seg_start,
                                                    int64_t *seg_len, int64_t *local, const int64_t capacity, int64_t &count,
                                                    int64_t &frames, int32_t &status)
{
    for (int base = 0; base < n; base += 32) {
        const int idx = base + lane;
        const int64_t slot = count + idx;
        const bool live = idx < n && slot < capacity;
        longlong2 e = make_longlong2(0, 0);
        if (idx < n) e = q[idx];
        const int64_t fr = live ? hubert_frames((int64_t)e.y) : 0;
        int64_t incl = fr;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (live) {
            seg_start[slot] = e.x;
            seg_len[slot] = e.y;
            if (local) local[slot] = frames + incl - fr;
        }
        frames += __shfl_sync(0xffffffffu, incl, 31);
    }
    count += n;
    if (count > capacity) status = AAT_ERR_CAPACITY;
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
    // optional fused frame CSR (every CTA writes its utterance's slice after a look-back)
    int64_t *seg_off;
    int64_t *n_seg;
    int64_t *utt_seg_off;
    int64_t *seg_local;   // [total_seg_slots] plan scratch: within-utterance frame offsets
    unsigned long long *utt_totals; // [n_utts] plan scratch, zero between launches: ready | frames << 32 | segments
    unsigned *ticket;
    int n_utts;
    int64_t min_frames, max_frames;
    int hop, n_mels, npts;
    float max_amp;
};

// Four-stage software pipeline over chunks of kChunk mel frames, one __syncthreads per iteration:
//     iteration `it`:   load chunk it      (worker warps)   amplitude curve -> s_amp[it & 1]
//                       scan chunk it - 1  (thread 0)        serial float32 cumsum -> s_cs ring
//                       test chunk it - 2  (worker warps)    running mean, local-max test, ordered compaction
//                       emit chunk it - 3  (thread 224)      merge/split state machine over the new boarders
// so the only thing on the critical path is the serial chain itself (~4-5 cycles per frame).
constexpr int kWorkers = 192;        // warps 1..6
constexpr int kEmitThread = 224;     // warp 7, lane 0
constexpr int kRing = 8192;          // cumsum ring (power of two) >= running_mean_points + 2 + 3 * kChunk

// -DAAT_POOL_TRACE (the trace build of profiles/): global-timer stamps of CTA 0..63, thread 0
#ifdef AAT_POOL_TRACE
__device__ unsigned long long g_bnd_trace[256 * 32];
#define BND_TRACE(slot)                                                                        \
    do {                                                                                       \
        if (threadIdx.x == 0 && blockIdx.x < 256 && (slot) < 32) {                             \
            unsigned long long t__;                                                            \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                            \
            g_bnd_trace[blockIdx.x * 32 + (slot)] = t__;                                       \
        }                                                                                      \
    } while (0)
#define BND_TRACE_ANY(slot)                                                                    \
    do {                                                                                       \
        if (blockIdx.x < 256 && (slot) < 32) {                                                 \
            unsigned long long t__;                                                            \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                            \
            g_bnd_trace[blockIdx.x * 32 + (slot)] = t__;                                       \
        }                                                                                      \
    } while (0)
#else
#define BND_TRACE(slot) \
    do {                \
    } while (0)
#define BND_TRACE_ANY(slot) \
    do {                    \
    } while (0)
#endif

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(kWorkers) : "memory"); }

AAT_TIMELINE_STORAGE(boundaries)
template <int kChunk>
__global__ void __launch_bounds__(kThreads) boundaries_kernel_t(const BoundaryParams p)
{
    AAT_TIMELINE_SCOPE(boundaries);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_amp = reinterpret_cast<float *>(smem_raw);          // [2][kChunk]
    float *s_cs = s_amp + 2 * kChunk;                             // [kRing]: cs index g lives at s_cs[g & (kRing - 1)]
    int *s_min = reinterpret_cast<int *>(s_cs + kRing);           // [2][kChunk] minima (global frame index) per chunk
    longlong2 *s_queue = reinterpret_cast<longlong2 *>(s_min + 2 * kChunk); // [kSegQueue] segments waiting to be written
    __shared__ int s_qctl[2]; // queue length, "chunk done" of the emit warp's current round
    __shared__ int s_wcount[kWorkers / 32];
    __shared__ int s_found[2];
    __shared__ int64_t s_total[2]; // this utterance's segment count and frame total (emit thread -> epilogue)

    BND_TRACE(0);
    const int tid = threadIdx.x;
    const int utt = blockIdx.x;
    pdl_wait(); // the log-mel kernel's mel / amp; also orders this kernel's writes after the previous readers
    pdl_launch_dependents();
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
    const int64_t n_chunks = (T + kChunk - 1) / kChunk;
    // candidate range [lo, hi) whose neighbourhood is complete once chunk c has been scanned
    auto cand_hi = [&](int64_t c) {
        const int64_t j1 = ((c + 1) * kChunk < T) ? (c + 1) * kChunk : T;
        int64_t hi = j1 - p.npts - 1; // needs cs[i + 1 + npts] < j1
        if (hi > L - 1) hi = L - 1;
        return hi < 1 ? (int64_t)1 : hi;
    };

    constexpr int kPre = (kChunk + kWorkers - 1) / kWorkers; // amplitude values each worker prefetches per chunk
    float pre[kPre];
    auto prefetch = [&](int64_t c) {
        const int64_t j0 = c * kChunk;
        const int w = tid - 32;
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int64_t t = j0 + w + k * kWorkers;
            pre[k] = (c < n_chunks && w + k * kWorkers < kChunk && t < T) ? amp_in[t] : 0.0f;
        }
    };
    if (amp_in && tid >= 32 && tid < 32 + kWorkers) prefetch(0);

    // merge/split state of the emit thread, in the narrowest index type that holds every quantity
    int64_t *local_ptr = p.seg_off ? p.seg_local + slot0 : nullptr;
    const int64_t kLim = 0x3fffffffLL;
    const bool narrow = n < kLim && p.min_frames < kLim && p.max_frames < kLim && capacity < kLim;
    MergeSplit<int32_t> ms32;
    MergeSplit<int64_t> ms64;
    int64_t seg_written = 0, seg_frames = 0; // emit warp: segments out so far and their HuBERT frames (warp-uniform)
    int32_t seg_status = 0;
    const int emit_lane = tid - kEmitThread; // 0..31 in the emit warp
    // One round of the emit warp: lane 0 decides (queue), the warp writes; repeated until the boarders are used up.
    auto emit_round = [&](const int found, auto next32, auto next64) {
        int i = 0;
        for (;;) {
            if (emit_lane == 0) {
                int nq;
                if (narrow)
                    nq = fill_segment_queue<int32_t>(ms32, i, found, next32, (int32_t)p.min_frames, (int32_t)p.max_frames, s_queue);
                else
                    nq = fill_segment_queue<int64_t>(ms64, i, found, next64, p.min_frames, p.max_frames, s_queue);
                s_qctl[0] = nq;
                s_qctl[1] = (i == found) && !(narrow ? ms32.splitting : ms64.splitting);
            }
            __syncwarp();
            const int nq = s_qctl[0];
            const int done = s_qctl[1];
            flush_segment_queue(s_queue, nq, emit_lane, seg_start, seg_len, local_ptr, capacity, seg_written, seg_frames,
                                seg_status);
            __syncwarp();
            if (done) break;
        }
    };
    int64_t n_minima = 0;                                                // workers and emit thread keep their own copy
    float run = 0.0f;                                                    // scan thread's carry

    BND_TRACE(1); // prologue done
    for (int64_t it = 0; it < n_chunks + 3; ++it) {
        if (tid == 0) {
            // ---------------- scan: chunk it - 1 ----------------
            const int64_t c = it - 1;
            if (c >= 0 && c < n_chunks) {
                if (c == 1) BND_TRACE_ANY(24); // scan of chunk 1 starts
                const int64_t j0 = c * kChunk;
                const int len = (int)(((j0 + kChunk < T) ? j0 + kChunk : T) - j0);
                const float *src_s = s_amp + (c & 1) * kChunk;
                float *dst_s = s_cs + (j0 & (kRing - 1));
                // 0.0f + x == x bit for bit (up to the sign of zero, which no later operation can see), so the
                // chain simply starts from run = 0 and numpy's cs[0] = x[0] needs no special case.
                const float4 *src = reinterpret_cast<const float4 *>(src_s);
                float4 *dst = reinterpret_cast<float4 *>(dst_s);
                const int nb = len / 16;
                float4 n0, n1, n2, n3;
                if (nb > 0) n0 = src[0], n1 = src[1], n2 = src[2], n3 = src[3];
#pragma unroll 2
                for (int bi = 0; bi < nb; ++bi) {
                    float4 a = n0, b = n1, c4 = n2, d = n3;
                    const int nx = (bi + 1 < nb) ? bi + 1 : bi; // next batch's loads fly while this batch's adds run
                    n0 = src[4 * nx], n1 = src[4 * nx + 1], n2 = src[4 * nx + 2], n3 = src[4 * nx + 3];
                    a.x = __fadd_rn(run, a.x);
                    a.y = __fadd_rn(a.x, a.y), a.z = __fadd_rn(a.y, a.z), a.w = __fadd_rn(a.z, a.w);
                    b.x = __fadd_rn(a.w, b.x), b.y = __fadd_rn(b.x, b.y), b.z = __fadd_rn(b.y, b.z), b.w = __fadd_rn(b.z, b.w);
                    c4.x = __fadd_rn(b.w, c4.x), c4.y = __fadd_rn(c4.x, c4.y), c4.z = __fadd_rn(c4.y, c4.z), c4.w = __fadd_rn(c4.z, c4.w);
                    d.x = __fadd_rn(c4.w, d.x), d.y = __fadd_rn(d.x, d.y), d.z = __fadd_rn(d.y, d.z), d.w = __fadd_rn(d.z, d.w);
                    dst[4 * bi] = a, dst[4 * bi + 1] = b, dst[4 * bi + 2] = c4, dst[4 * bi + 3] = d;
                    run = d.w;
                }
                for (int i = nb * 16; i < len; ++i) {
                    run = __fadd_rn(run, src_s[i]);
                    dst_s[i] = run;
                }
                if (c == 1) BND_TRACE_ANY(25); // ... ends
            }
        } else if (tid >= 32 && tid < 32 + kWorkers) {
            const int w = tid - 32;
            // ---------------- test: chunk it - 2 ----------------
            {
                const int64_t c = it - 2;
                if (c >= 0 && c < n_chunks) {
                    if (c == 1 && w == 0) BND_TRACE_ANY(26); // test of chunk 1 starts
                    const int64_t lo = (c == 0) ? 1 : cand_hi(c - 1);
                    const int64_t hi = cand_hi(c);
                    const int64_t range = hi > lo ? hi - lo : 0;
                    const int per = (int)((range + kWorkers - 1) / kWorkers); // <= 3 (< 32 mask bits)
                    const int64_t a = lo + (int64_t)w * per;
                    unsigned mask = 0;
                    if (per > 0 && a < hi) {
                        auto rm = [&](int64_t i) {
                            return __fdiv_rn(__fsub_rn(s_cs[(i + p.npts) & (kRing - 1)], s_cs[i & (kRing - 1)]), nf);
                        };
                        float rl = rm(a - 1), rc = rm(a);
                        for (int u = 0; u < per; ++u) {
                            const int64_t i = a + u;
                            if (i >= hi) break;
                            const float rr = rm(i + 1);
                            if (rc > __fadd_rn(rr, 1e-5f) && rc > __fadd_rn(rl, 1e-5f) && rc > p.max_amp) mask |= 1u << u;
                            rl = rc, rc = rr;
                        }
                    }
                    // ordered compaction over the 192 workers
                    const int cnt = __popc(mask);
                    int incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, d);
                        if ((w & 31) >= d) incl += v;
                    }
                    if ((w & 31) == 31) s_wcount[w >> 5] = incl;
                    worker_barrier();
                    int base = 0, total = 0;
#pragma unroll
                    for (int k = 0; k < kWorkers / 32; ++k) {
                        const int v = s_wcount[k];
                        if (k < (w >> 5)) base += v;
                        total += v;
                    }
                    int *mq = s_min + (c & 1) * kChunk;
                    int pos = base + incl - cnt;
                    unsigned m = mask;
                    while (m) {
                        const int u = __ffs(m) - 1;
                        m &= m - 1;
                        mq[pos] = (int)(a + u);
                        if (minima_out) minima_out[n_minima + pos] = a + u;
                        ++pos;
                    }
                    if (w == 0) s_found[c & 1] = total;
                    n_minima += total;
                    worker_barrier(); // s_wcount is reused next iteration
                    if (c == 1 && w == 0) BND_TRACE_ANY(27); // ... ends
                }
            }
            // ---------------- load: chunk it ----------------
            // With a precomputed amplitude curve the values were fetched into registers one iteration ago (so
            // the global-memory latency hides behind a whole pipeline iteration); the fetch for chunk it + 1
            // is issued now and consumed next time round.
            {
                const int64_t c = it;
                if (c < n_chunks) {
                    const int64_t j0 = c * kChunk;
                    const int len = (int)(((j0 + kChunk < T) ? j0 + kChunk : T) - j0);
                    float *dst = s_amp + (c & 1) * kChunk;
                    if (amp_in) {
#pragma unroll
                        for (int k = 0; k < kPre; ++k) {
                            const int i = w + k * kWorkers;
                            if (i < len) dst[i] = pre[k];
                        }
                    } else { // numpy mean(axis=0): rows added in order, float32 (ref:src/aat/tokenizer.py:67)
                        for (int i = w; i < len; i += kWorkers) {
                            const int64_t t = j0 + i;
                            float acc = mel[t];
#pragma unroll 8
                            for (int r = 1; r < p.n_mels; ++r) acc = __fadd_rn(acc, mel[(size_t)r * T + t]);
                            dst[i] = __fmul_rn(-10.0f, __fdiv_rn(acc, (float)p.n_mels));
                        }
                    }
                }
                if (amp_in) prefetch(it + 1);
                if (it == 3 && w == 0) BND_TRACE_ANY(9); // workers done with iteration 3 (test chunk 1, load chunk 3)
            }
        } else if (tid >= kEmitThread) {
            // ---------------- emit: chunk it - 3 (warp 7) ----------------
            const int64_t c = it - 3;
            if (c >= 0 && c < n_chunks) {
                if (emit_lane == 0) BND_TRACE_ANY(10 + 2 * (int)c); // merge/split: start of chunk c
                const int found = s_found[c & 1];
                const int *mq = s_min + (c & 1) * kChunk;
                emit_round(found, [&](int i) { return (int32_t)(mq[i] * p.hop); },
                           [&](int i) { return (int64_t)mq[i] * p.hop; });
                n_minima += found;
                if (emit_lane == 0) BND_TRACE_ANY(11 + 2 * (int)c); // ... end of chunk c
            }
        }
        __syncthreads();
        BND_TRACE(2 + (int)it); // end of pipeline iteration `it`
    }

    if (tid >= kEmitThread) {
        // the end of the waveform is the last boarder (ref:src/aat/tokenizer.py:137), then the zero-padded tail of
        // min_segment_frames samples (ref:src/aat/tokenizer.py:177-181)
        if (emit_lane == 0) BND_TRACE_ANY(20);
        emit_round(1, [&](int) { return (int32_t)n; }, [&](int) { return n; });
        if (emit_lane == 0) BND_TRACE_ANY(21);
        // only lane 0 ran the merge/split decisions: the last accepted boarder lives in its registers
        const int64_t prev = __shfl_sync(0xffffffffu, narrow ? (int64_t)ms32.prev : ms64.prev, 0);
        if (prev != n) {
            if (seg_status == 0) seg_status = (n - prev > p.min_frames) ? AAT_ERR_TAIL : 1;
            if (emit_lane == 0) s_queue[0] = make_longlong2((long long)prev, (long long)p.min_frames);
            __syncwarp();
            flush_segment_queue(s_queue, 1, emit_lane, seg_start, seg_len, local_ptr, capacity, seg_written, seg_frames,
                                seg_status);
        }
    }
    if (tid == kEmitThread) {
        BND_TRACE_ANY(22);
        const int64_t count = seg_written, frames = seg_frames;
        const int32_t status = seg_status;
        p.seg_count[utt] = (int32_t)(count < capacity ? count : capacity);
        if (p.minima_count) p.minima_count[utt] = (int32_t)n_minima;
        p.status[utt] = status;
        if (p.seg_off) {
            s_total[0] = count < capacity ? count : capacity, s_total[1] = frames;
            // one 64-bit store: ready bit | frames | segment count.  The word carries everything the later utterances'
            // look-back needs, so neither side needs release/acquire ordering (no membar, no L1 invalidation).
            __stcg(p.utt_totals + utt, kTotalsReady | ((unsigned long long)frames << 32) |
                                           (unsigned long long)(uint32_t)s_total[0]);
        }
    }

    // ---- fused epilogue: the packed frame CSR by decoupled look-back ----
    // Every utterance's CTA publishes (segments, frames) in one word, sums the words of the utterances before it
    // (they are resident or done: CTAs are dispatched in index order, and nobody waits before publishing) and
    // writes its own slice of seg_off.  No CTA does serial work for the others: the epilogue costs one round trip
    // after the slowest utterance instead of a rebase pass by the last CTA (6 us at 64 utterances,
    // profiles/r1_bnd_timeline.txt).  The last CTA through the ticket clears the words for the next launch.
    if (tid == kEmitThread) BND_TRACE_ANY(23);
    if (p.seg_off != nullptr) {
        __shared__ int64_t s_wsum[2][kThreads / 32];
        __shared__ int s_last;
        int64_t cnt = 0, frm = 0;
        for (int b = tid; b < utt; b += kThreads) {
            unsigned long long v;
            // back off between polls: thousands of threads spinning on the same few L2 lines delay the very stores they
            // are waiting for (the last word took 3.8 us to be seen without it)
            while (!((v = ld_relaxed_u64(p.utt_totals + b)) & kTotalsReady)) __nanosleep(100);
            cnt += (int64_t)(v & 0xffffffffull);
            frm += (int64_t)((v >> 32) & 0x7fffffffull);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            frm += __shfl_xor_sync(0xffffffffu, frm, d);
        }
        if ((tid & 31) == 0) s_wsum[0][tid >> 5] = cnt, s_wsum[1][tid >> 5] = frm;
        __syncthreads(); // also: the emit thread's s_total and its seg_local stores are visible to the CTA
        BND_TRACE(28); // look-back done
        int64_t base_seg = 0, base_frm = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) base_seg += s_wsum[0][w], base_frm += s_wsum[1][w];
        const int64_t my_cnt = s_total[0], my_frm = s_total[1];
        const int64_t *local = p.seg_local + slot0;
        int64_t *dst = p.seg_off + base_seg;
        for (int64_t i = tid; i < my_cnt; i += kThreads) dst[i] = base_frm + __ldcg(reinterpret_cast<const long long *>(local + i));
        if (tid == 0) {
            if (p.utt_seg_off) p.utt_seg_off[utt] = base_seg;
            if (utt == p.n_utts - 1) {
                if (p.utt_seg_off) p.utt_seg_off[p.n_utts] = base_seg + my_cnt;
                p.n_seg[0] = base_seg + my_cnt;
                p.n_seg[1] = base_frm + my_frm; // the frames the CSR covers (AAT_POOL_ROWS_FROM_DEVICE)
                p.seg_off[base_seg + my_cnt] = base_frm + my_frm;
            }
            s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1); // after this CTA's look-back
        }
        __syncthreads();
        if (s_last) { // everybody has read every word: clear them for the next launch (graph replays included)
            for (int b = tid; b < p.n_utts; b += kThreads) p.utt_totals[b] = 0ull;
            if (tid == 0) *p.ticket = 0;
        }
        BND_TRACE(29);
    }
}

#ifdef AAT_POOL_TRACE
} // namespace
extern "C" __attribute__((visibility("default"))) int aat_debug_bnd_trace(unsigned long long *out_host, int n_ctas)
{
    return (int)cudaMemcpyFromSymbol(out_host, g_bnd_trace, sizeof(unsigned long long) * 32 * (size_t)n_ctas);
}
namespace {
#endif

// Stand-alone amplitude curve: amp[t] = -10 * mean_r(mel[r][t]) with numpy's arithmetic (rows added in order in
// float32, exact division by the row count; ref:src/aat/tokenizer.py:67).  One thread per frame, consecutive threads on
// consecutive frames (every row access of a warp is one 128-byte line), the loads of 16 rows in flight before the
// dependent additions consume them.  This is the log-mel kernel's fused epilogue as a separate, fully parallel pass:
// the pipelined schedule (aat_b200/pipeline.py) prefers it, because there the extra pass over the mel (L2-resident at
// batch sizes up to ~100 MB of mel) hides behind the next batch's log-mel, while the epilogue inside the log-mel kernel
// costs that kernel a CTA barrier and a serial 64-term chain per tile (5 % of its time).
__global__ void __launch_bounds__(256)
amplitude_kernel(const float *__restrict__ mel, float *__restrict__ amp, const int64_t *n_samples, const int64_t *frame_off,
                 int hop, int n_mels)
{
    pdl_wait(); // the log-mel kernel's mel
    pdl_launch_dependents();
    const int utt = blockIdx.y;
    const int64_t T = 1 + n_samples[utt] / hop;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t fbase = frame_off[utt];
    const float *col = mel + (size_t)n_mels * fbase + t;
    float acc = col[0];
    int r = 1;
    for (; r + 16 <= n_mels; r += 16) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __ldg(col + (size_t)(r + k) * T);
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = __fadd_rn(acc, v[k]);
    }
    for (; r < n_mels; ++r) acc = __fadd_rn(acc, __ldg(col + (size_t)r * T));
    amp[fbase + t] = __fmul_rn(-10.0f, __fdiv_rn(acc, (float)n_mels));
}

__global__ void process_boarders_kernel(int64_t n_samples, const int64_t *boarders, int64_t n_boarders,
                                        int64_t min_frames, int64_t max_frames, int64_t *seg_start, int64_t *seg_len,
                                        int64_t capacity, int32_t *seg_count, int32_t *status)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    SegState st{0, 0, 0, 0, nullptr}; // caller-supplied boarders may be anything: keep the 64-bit machine here
    for (int64_t i = 0; i < n_boarders; ++i)
        push_boarder<int64_t>(boarders[i], min_frames, max_frames, seg_start, seg_len, capacity, st);
    finish_segments<int64_t>(n_samples, min_frames, seg_start, seg_len, capacity, st);
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
    p.ticket = reinterpret_cast<unsigned *>(plan->d_mel_sched + 2); // per plan, so plans on different streams are independent
    // Stage size: 128-frame stages were measured slower (31 us vs 27 us at 64 x 16 s) and 256 equal: the kernel is
    // bound by the merge/split thread, not by the pipeline fill (profiles/r1_bnd_timeline.txt).
    constexpr int chunk = kChunk;
    const bool fuse_csr = seg_off != nullptr; // look-back epilogue: no scratch beyond one word per utterance
    p.seg_off = fuse_csr ? seg_off : nullptr;
    p.n_seg = n_seg;
    p.utt_seg_off = utt_seg_off;
    p.seg_local = plan->d_seg_local;
    p.utt_totals = reinterpret_cast<unsigned long long *>(plan->d_utt_frames);
    const size_t smem = sizeof(float) * (size_t)(2 * chunk + kRing) + sizeof(int) * (size_t)(2 * chunk) +
                        sizeof(longlong2) * (size_t)kSegQueue;
    auto kernel = boundaries_kernel_t<kChunk>;
    AAT_CUDA_CHECK(prepare_kernel(ctx, kernel, 0, smem, nullptr));
    {
        ProfileScope prof(ctx, AAT_K_BOUNDARIES, stream);
        AAT_CUDA_CHECK(launch_pdl(kernel, dim3(plan->n_utts), dim3(kThreads), smem, stream, p));
        AAT_LAUNCH_CHECK();
    }
    if (seg_off != nullptr && !fuse_csr) // batch too large for the fused epilogue's scratch: separate kernel
        return launch_segment_frame_csr(ctx, plan, seg_len, seg_count, seg_off, n_seg, utt_seg_off, stream);
    return AAT_OK;
}

int launch_amplitude(aat_ctx *ctx, const aat_plan *plan, const float *mel, float *amp, cudaStream_t stream)
{
    if (plan->n_utts == 0 || plan->max_frames == 0) return AAT_OK;
    AAT_MAX_SMEM_CARVEOUT(amplitude_kernel);
    const dim3 grid((unsigned)((plan->max_frames + 255) / 256), (unsigned)plan->n_utts);
    AAT_CUDA_CHECK(launch_pdl(amplitude_kernel, grid, dim3(256), 0, stream, mel, amp, (const int64_t *)plan->d_n_samples,
                              (const int64_t *)plan->d_frame_off, (int)ctx->cfg.hop_length, (int)ctx->cfg.num_mel_filters));
    AAT_LAUNCH_CHECK();
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
    AAT_CUDA_CHECK(prepare_kernel(ctx, segment_frame_csr_kernel, 0, smem, nullptr));
    ProfileScope prof(ctx, AAT_K_FRAME_CSR, stream);
    segment_frame_csr_kernel<<<1, kCsrThreads, smem, stream>>>(plan->n_utts, plan->d_seg_slot_off, seg_len, seg_count,
                                                               seg_off, n_seg, utt_seg_off);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat

AAT_TIMELINE_EXPORT(boundaries, aat::)
