// K4: ragged per-segment mean-pool of frame embeddings — the HBM-streaming kernel.
//
// Replaces `torch.cat([x.mean(dim=1, keepdim=True).to(float32) for x in embs], dim=1)`
// (ref:scripts/mean_hubert_embeddings.py:19-20) on the packed layout emb[n_rows, dim] + CSR
// seg_off[S+1].  Pure streaming reduction (~0.25 flop/B): the only roofline is HBM bandwidth, so the
// design is about bytes in flight and balance, not arithmetic:
//
//   * work is tiled by BYTES, not by segment: the row range [0, n_rows) is split evenly over a
//     persistent grid (CTAs-per-SM x 148 SMs), so 6-frame and 74-frame segments cost the same per byte
//     and a giant segment cannot serialise on one SM;
//   * each CTA's rows are one contiguous byte range, streamed once through a ring of shared-memory
//     stages by 1-D bulk async copies (cp.async.bulk.shared.global -> SASS UBLKCP) issued by a single
//     producer thread and tracked by mbarrier transaction counts — tens of KB in flight per SM without
//     spending registers or address arithmetic on it;
//   * consumer threads own one 16-byte column slab each (128-bit conflict-free shared loads), walk the
//     rows of a stage, keep four interleaved float32 accumulators and flush at every segment boundary
//     with a coalesced float4 store of sum / n;
//   * a segment cut by a CTA boundary is finished deterministically: the CTA that holds its first row
//     owns it; the CTAs that hold its middle and end publish their partial sums (+ release flag) as
//     soon as they have them and never wait first; the owner adds them in CTA order at the very end
//     of its own rows.  No CTA waits on a waiter and the whole grid is co-resident, so there is neither
//     a dependency chain nor a deadlock; flags are reset by their single consumer (graph-replay safe);
//   * optional epilogue: float64 column sums of the pooled vectors (input of the dataset-mean
//     allreduce) are accumulated per CTA and reduced in CTA order by a second tiny kernel;
//   * launched with programmatic dependent launch.  When the caller vouches that the kernel in front of this one on
//     the stream does not write the embeddings (AAT_POOL_EMB_READY: the library's own boundary scan, whose outputs
//     are only the offsets), the whole ring of embedding stages is requested before the kernel waits for that
//     predecessor; otherwise nothing is read before the wait.  The first offsets window arrives by one bulk copy, and
//     the owner of a cut segment looks at its neighbour's partial sums two stages before its last row
//     (profiles/r1_pool_timeline*.txt, r1_pool_ab.txt);
//   * the cross-CTA scratch (partial sums, flags, per-CTA column sums) belongs to the plan the launch names, or to the
//     context for launches without a plan, which are ordered against each other by an event: two launches never share
//     a scratch block while in flight.  An owner only ever waits for CTAs with a HIGHER index, which publish without
//     waiting for anything, so with CTAs dispatched in index order the kernel makes progress whatever else occupies
//     the SMs (a second pool launch on another stream, another tenant).
//
// Algorithmic bytes per launch: n_rows*dim*e + S*dim*4 + (S+1)*8  (SURVEY.md §8d).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "aat_internal.cuh"
#include <cstdlib>

namespace aat {

namespace {

// -DAAT_POOL_TRACE (profiles/pool_timeline.py builds a separate library with it): every CTA records the global
// timer at six points of its life; never compiled into the product library.
#ifdef AAT_POOL_TRACE
__device__ unsigned long long g_pool_trace[1024 * 8];
__device__ __forceinline__ unsigned long long trace_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define POOL_TRACE(slot)                                                     \
    do {                                                                     \
        if (threadIdx.x == 0) g_pool_trace[blockIdx.x * 8 + (slot)] = trace_now(); \
    } while (0)
#else
#define POOL_TRACE(slot) \
    do {                 \
    } while (0)
#endif

// Ring geometry.  Every stage costs a fixed ~0.45 us of hand-shaking (consumers release it, the producer thread sees the
// release, issues the bulk copy, the data crosses the chip, the consumers see it land), so FEW, BIG stages win: two
// stages of 48 KB instead of four of 24 KB in the same 96 KB stream 5-7 % faster at every size (config 2: 29.7 ->
// 28.0 us back to back, configs 3 / 4: 167 -> 161 / 350 -> 336 us; profiles/r2_pool_ring_grid.txt), and many small stages
// are far worse (7 x 9 KB: 72 us).
#ifndef AAT_POOL_STAGES
#define AAT_POOL_STAGES 2
#define AAT_POOL_STAGE_KB 48
#define AAT_POOL_CTAS 2
#endif
constexpr int kStages = AAT_POOL_STAGES;
constexpr int kStageBytes = AAT_POOL_STAGE_KB * 1024; // 48 KB: 16 rows of 768 fp32, 12 rows of 1024 fp32
constexpr int kMaxConsumers = 256;
constexpr int kMaxSlabs = 4; // 16-byte column slabs per consumer thread -> dim*e <= 16 KB
constexpr int kMaxCtasPerSm = AAT_POOL_CTAS;
// Stages requested before the dependency wait when the caller allows it (AAT_POOL_EMB_READY); measured at config 2:
// 1 -> 0.1907 ms/step, 2 -> 0.1899, 4 -> 0.1895, and the event-timed kernel alone is no slower (gpurun b18).
constexpr int kPreStages = kStages;
constexpr int kOffCache = 256; // segment offsets of the CTA's neighbourhood kept in shared memory
// A segment that spans many CTAs (ragged stress: one giant segment) is reduced in two levels: the first CTA of every
// aligned group of kGroup CTAs that lie wholly inside the segment adds its group's pieces and publishes ONE group
// piece, and the owner adds group pieces.  With one level the owner fetched ~295 pieces, 8 per round trip (+75 us at
// config-3 size, profiles/r2_pool_sweep.txt); now it is <= 2 + (296 / 16 + 2 * 15) / 8 round trips.
constexpr int kGroup = 16;

// ---------------------------------------------------------------- PTX helpers (mbarrier + bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_barrier(int n_consumers)
{
    asm volatile("bar.sync 1, %0;" ::"r"(n_consumers) : "memory");
}
// barrier over the consumer threads that also ANDs a predicate across them
__device__ __forceinline__ bool consumer_barrier_and(int n_consumers, bool pred)
{
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 q, %2, 0;\n"
        "bar.red.and.pred p, 1, %1, q;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(r)
        : "r"(n_consumers), "r"((uint32_t)pred)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------- element traits: one 16-byte slab
template <typename T>
struct Slab;
template <>
struct Slab<float> {
    static constexpr int kCols = 4;
    __device__ static void load(const void *p, float (&v)[4])
    {
        const float4 x = *reinterpret_cast<const float4 *>(p);
        v[0] = x.x, v[1] = x.y, v[2] = x.z, v[3] = x.w;
    }
    // torch: sum.div_(n) in float32
    __device__ static float finish(float sum, float n) { return __fdiv_rn(sum, n); }
};
template <>
struct Slab<__half> {
    static constexpr int kCols = 8;
    __device__ static void load(const void *p, float (&v)[8])
    {
        const uint4 x = *reinterpret_cast<const uint4 *>(p);
        const __half2 *h = reinterpret_cast<const __half2 *>(&x);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            v[2 * i] = f.x, v[2 * i + 1] = f.y;
        }
    }
    // torch: mean of a half tensor accumulates and divides in float32 and rounds once to half
    // (verified against torch 2.11 CPU: half(sum_f32 / n) reproduces x.mean(dim=1) bit for bit)
    __device__ static float finish(float sum, float n) { return __half2float(__float2half_rn(__fdiv_rn(sum, n))); }
};
template <>
struct Slab<__nv_bfloat16> {
    static constexpr int kCols = 8;
    __device__ static void load(const void *p, float (&v)[8])
    {
        const uint4 x = *reinterpret_cast<const uint4 *>(p);
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static float finish(float sum, float n)
    {
        return __bfloat162float(__float2bfloat16_rn(__fdiv_rn(sum, n)));
    }
};

struct PoolParams {
    const unsigned char *emb;
    const int64_t *seg_off;
    const int64_t *n_seg_dev;
    float *out;
    float *head;      // [G, dim] this CTA's piece of a segment that began in an earlier CTA
    int *head_flag;   // [G]
    double *colsum;   // [G, dim] per-CTA column sums of pooled vectors (optional)
    int64_t n_rows;
    int64_t n_seg;
    int dim;
    int row_bytes;
    int rows_per_stage;
    int slabs_per_row; // row_bytes / 16
    int n_consumers;
    int group_base;    // first row of head / head_flag that holds group pieces (= the scratch block's max_ctas)
    int pre_stages;    // stages requested before the dependency wait (0 unless the caller set AAT_POOL_EMB_READY)
    int rows_from_dev; // n_seg_dev[1] holds the number of rows the CSR covers; n_rows is an upper bound
};

__device__ __forceinline__ int64_t cta_row_begin(int64_t c, int64_t n_rows, int64_t G) { return (c * n_rows) / G; }

// Number of entries of off[0..count) that are <= key, found cooperatively by the consumer threads.
__device__ int64_t coop_upper_bound(const int64_t *off, int64_t count, int64_t key, int tid, int n_consumers,
                                    int *s_votes)
{
    int64_t lo = 0, hi = count; // entries < lo are <= key; entries >= hi are > key
    while (lo < hi) {
        const int64_t span = hi - lo;
        const int64_t step = (span + n_consumers - 1) / n_consumers;
        const int64_t q = lo + (int64_t)(tid + 1) * step - 1;
        const bool pred = (q < hi) && (off[q] <= key);
        const unsigned ballot = __ballot_sync(0xffffffffu, pred);
        if ((tid & 31) == 0) s_votes[tid >> 5] = __popc(ballot);
        consumer_barrier(n_consumers);
        int c = 0;
        for (int w = 0; w < n_consumers / 32; ++w) c += s_votes[w];
        consumer_barrier(n_consumers);
        const int64_t new_lo = lo + (int64_t)c * step;
        int64_t new_hi = lo + (int64_t)(c + 1) * step - 1;
        if (new_hi > hi) new_hi = hi;
        lo = new_lo < hi ? new_lo : hi;
        hi = new_hi;
    }
    return lo;
}

// First window of segment offsets a CTA looks at: centred on the segment index that a uniform segment length
// would give for row r0.  When the source is 16-byte aligned the producer fetches it with ONE bulk copy issued
// ahead of the embedding stream — a plain load issued a microsecond later queues behind ~28 MB of bulk traffic
// from all CTAs and takes 3.5-6 us (profiles/r1_pool_timeline.txt).
struct OffWindow {
    int64_t first;
    int count;
    bool bulk;
};
__device__ __forceinline__ OffWindow first_window(const int64_t *seg_off, int64_t r0, int64_t n_rows, int64_t S_total)
{
    OffWindow w;
    const int64_t guess = (int64_t)((double)r0 / (double)n_rows * (double)S_total);
    int64_t w0 = guess - kOffCache / 2;
    const int64_t w_max = S_total + 1 - kOffCache;
    if (w0 > w_max) w0 = w_max;
    if (w0 < 0) w0 = 0;
    if ((reinterpret_cast<uintptr_t>(seg_off + w0) & 15) != 0 && w0 > 0) --w0;
    const int64_t avail = S_total + 1 - w0;
    w.first = w0;
    w.count = (int)(avail < kOffCache ? avail : kOffCache);
    w.bulk = (reinterpret_cast<uintptr_t>(seg_off + w0) & 15) == 0 && w.count >= 2;
    if (w.bulk) w.count &= ~1; // whole 16-byte units
    return w;
}

AAT_TIMELINE_STORAGE(pool)
template <typename EmbT, int kSlabs, bool kColsum>
__global__ void __launch_bounds__(kMaxConsumers + 32) pool_kernel(const PoolParams p)
{
    AAT_TIMELINE_SCOPE(pool);
    using S = Slab<EmbT>;
    constexpr int kCols = S::kCols;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kStages];
    __shared__ __align__(8) uint64_t s_empty[kStages];
    __shared__ __align__(8) uint64_t s_offbar;
    __shared__ int s_votes[kMaxConsumers / 32];
    __shared__ __align__(16) int64_t s_off[kOffCache];

    POOL_TRACE(0); // CTA entry
    const int tid = threadIdx.x;
    const int n_consumers = p.n_consumers;
    const int64_t G = gridDim.x;
    const int64_t c = blockIdx.x;
    int64_t n_rows = p.n_rows;
    int64_t r0 = cta_row_begin(c, n_rows, G);
    int64_t r1 = cta_row_begin(c + 1, n_rows, G);
    int64_t n_chunks = (r1 - r0 + p.rows_per_stage - 1) / p.rows_per_stage;
    const size_t stage_stride = (size_t)p.rows_per_stage * p.row_bytes;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], n_consumers / 32);
        }
        mbar_init(&s_offbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_stage = [&](int64_t ch) {
        const int s = (int)(ch % kStages);
        const int64_t row = r0 + ch * p.rows_per_stage;
        const int64_t rows = (r1 - row < p.rows_per_stage) ? r1 - row : p.rows_per_stage;
        const uint32_t bytes = (uint32_t)(rows * p.row_bytes);
        mbar_expect_tx(&s_full[s], bytes);
        bulk_g2s(smem_raw + s * stage_stride, p.emb + (size_t)row * p.row_bytes, bytes, &s_full[s]);
    };
    // When the caller vouches that the embedding rows are not written by the kernel in front of this one (the
    // library's boundary scan), the whole ring is requested before the dependency wait: under programmatic dependent
    // launch it lands while that kernel is still running.  Segment offsets and counts are its outputs and are only
    // touched after the wait.
    const bool producer = tid == n_consumers;
    const int64_t n_pre = n_chunks < p.pre_stages ? n_chunks : p.pre_stages;
    if (producer)
        for (int64_t ch = 0; ch < n_pre; ++ch) issue_stage(ch);
    pdl_wait();
    pdl_launch_dependents();
    const int64_t S_total = p.n_seg_dev ? min(p.n_seg_dev[0], p.n_seg) : p.n_seg;
    if (p.rows_from_dev) { // the rows the CSR covers, as counted on the device: nothing beyond them is streamed
        const int64_t covered = p.n_seg_dev[1];
        if (covered < n_rows) n_rows = covered < 0 ? 0 : covered;
        r0 = cta_row_begin(c, n_rows, G);
        r1 = cta_row_begin(c + 1, n_rows, G);
        n_chunks = (r1 - r0 + p.rows_per_stage - 1) / p.rows_per_stage;
    }

    if (tid >= n_consumers) {
        // ============================== producer warp ==============================
        const int lane = tid - n_consumers;
        if (lane == 0 && r0 < r1 && S_total <= 0) {
            for (int64_t ch = 0; ch < n_pre; ++ch) mbar_wait(&s_full[ch], 0); // nothing to pool: let the speculative stages land
        } else if (lane == 0 && r0 < r1) {
            // the offsets window goes ahead of the bulk of the stream: a plain load issued later would queue behind
            // ~28 MB of bulk traffic from all CTAs and take 3.5-6 us (profiles/r1_pool_timeline_before.txt)
            const OffWindow w = first_window(p.seg_off, r0, n_rows, S_total);
            if (w.bulk) {
                mbar_expect_tx(&s_offbar, (uint32_t)(w.count * sizeof(int64_t)));
                bulk_g2s(s_off, p.seg_off + w.first, (uint32_t)(w.count * sizeof(int64_t)), &s_offbar);
            }
            for (int64_t ch = n_pre; ch < n_chunks; ++ch) {
                const int64_t use = ch / kStages;
                if (use > 0) mbar_wait(&s_empty[(int)(ch % kStages)], (uint32_t)((use - 1) & 1));
                issue_stage(ch);
            }
        } else if (lane != 0) {
            // idle lanes: empty segments never meet a row, so they are written here (torch: mean of
            // an empty slice is NaN).  Grid-stride over all segments, 31 lanes per CTA.
            const float nan = __int_as_float(0x7fc00000);
            for (int64_t s = c * 31 + (lane - 1); s < S_total; s += G * 31) {
                if (p.seg_off[s] == p.seg_off[s + 1]) {
                    float *o = p.out + (size_t)s * p.dim;
                    for (int d = 0; d < p.dim; ++d) o[d] = nan;
                }
            }
        }
        return;
    }

    // ================================ consumers ================================
    float acc[kSlabs][4][kCols];
    double csum[kColsum ? kSlabs : 1][kCols];
#pragma unroll
    for (int j = 0; j < kSlabs; ++j)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < kCols; ++k) acc[j][u][k] = 0.0f;
    if (kColsum) {
#pragma unroll
        for (int j = 0; j < kSlabs; ++j)
#pragma unroll
            for (int k = 0; k < kCols; ++k) csum[j][k] = 0.0;
    }

    if (r0 < r1 && S_total > 0) {
        // current segment: last s with seg_off[s] <= r0.  Segments have a typical length, so the answer is
        // near r0 * S / n_rows: load a window of offsets around that guess into shared memory (ONE global
        // round trip, and the window doubles as the boundary cache) and finish with a binary search in shared
        // memory; only if the guess misses fall back to the cooperative k-ary search over global memory.
        int64_t cache_base = 0;
        int cache_n = 0;
        auto fill_cache = [&](int64_t first) {
            consumer_barrier(n_consumers); // nobody still reads the old window
            cache_base = first;
            const int64_t avail = S_total + 1 - first;
            cache_n = (int)(avail < kOffCache ? avail : kOffCache);
            for (int i = tid; i < cache_n; i += n_consumers) s_off[i] = p.seg_off[first + i];
            consumer_barrier(n_consumers);
        };
        int64_t idx;
        {
            const OffWindow w = first_window(p.seg_off, r0, n_rows, S_total);
            if (w.bulk) {
                mbar_wait(&s_offbar, 0); // the producer's bulk copy of the window has landed
                cache_base = w.first;
                cache_n = w.count;
            } else {
                fill_cache(w.first);
            }
            const int64_t w0 = w.first;
            const bool lower_ok = w0 == 0 || s_off[0] <= r0;
            const bool upper_ok = w0 + cache_n == S_total + 1 || s_off[cache_n - 1] > r0;
            if (lower_ok && upper_ok) {
                int lo = 0, hi = cache_n; // entries of the window that are <= r0
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_off[mid] <= r0) lo = mid + 1; else hi = mid;
                }
                idx = w0 + lo;
            } else {
                idx = coop_upper_bound(p.seg_off, S_total + 1, r0, tid, n_consumers, s_votes);
            }
        }
        int64_t seg = idx - 1; // -1: rows before the first segment; S_total: rows after the last one
        int64_t seg_begin, seg_end;
        bool in_gap;
        auto load_segment = [&]() {
            if (seg < 0) {
                if (cache_n == 0 || cache_base != 0) fill_cache(0);
                in_gap = true, seg_begin = r0, seg_end = s_off[0];
            } else if (seg >= S_total) {
                in_gap = true, seg_begin = r0, seg_end = INT64_MAX;
            } else {
                if (seg < cache_base || seg + 1 >= cache_base + cache_n) fill_cache(seg);
                in_gap = false, seg_begin = s_off[seg - cache_base], seg_end = s_off[seg + 1 - cache_base];
            }
        };
        load_segment();
        POOL_TRACE(1); // first segment located
        auto reduce_acc = [&](int j, int k) { return (acc[j][0][k] + acc[j][1][k]) + (acc[j][2][k] + acc[j][3][k]); };
        auto clear_acc = [&]() {
#pragma unroll
            for (int j = 0; j < kSlabs; ++j)
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int k = 0; k < kCols; ++k) acc[j][u][k] = 0.0f;
        };
        auto write_pooled = [&](int j, const float (&sum)[kCols], float nrows) {
            const int slab = tid + j * n_consumers;
            if (slab < p.slabs_per_row) {
                float *o = p.out + (size_t)seg * p.dim + (size_t)slab * kCols;
                float r[kCols];
#pragma unroll
                for (int k = 0; k < kCols; ++k) {
                    r[k] = S::finish(sum[k], nrows);
                    if (kColsum) csum[kColsum ? j : 0][k] += (double)r[k];
                }
#pragma unroll
                for (int k = 0; k < kCols; k += 4)
                    *reinterpret_cast<float4 *>(o + k) = make_float4(r[k], r[k + 1], r[k + 2], r[k + 3]);
            }
        };

        // Early look at the next CTA's piece: if this CTA turns out to own a segment that runs on into CTA c + 1
        // only, the final flush finds flag and partial sums already in registers instead of paying two dependent
        // global round trips (~2 us under load, on every CTA's critical path) after its last row.
        constexpr bool kEarlyCarry = kSlabs == 1;
        int carry_flag = 0;
        float carry[kCols];
        auto prefetch_carry = [&]() {
            if (kEarlyCarry && tid >= p.slabs_per_row) carry_flag = 1; // no slab: neutral in the vote
            if (kEarlyCarry && c + 1 < G && tid < p.slabs_per_row) {
                carry_flag = ld_acquire(p.head_flag + c + 1);
                if (carry_flag) {
#pragma unroll
                    for (int k = 0; k < kCols; ++k) carry[k] = __ldcg(p.head + (size_t)(c + 1) * p.dim + tid * kCols + k);
                }
            }
        };

        // last CTA that holds rows of a segment ending at seg_end
        auto last_cta_of = [&](int64_t seg_end_) {
            const int64_t last_row = seg_end_ - 1;
            int64_t ce = n_rows > 0 ? (last_row < n_rows ? (last_row * G) / n_rows : G - 1) : G - 1;
            while (ce + 1 < G && cta_row_begin(ce + 1, n_rows, G) <= last_row) ++ce;
            while (ce > c && cta_row_begin(ce, n_rows, G) > last_row) --ce;
            return ce;
        };
        const bool every_cta_has_rows = n_rows >= G; // both sides of the group protocol evaluate the same condition
        // sum[j][k] += pieces lo..hi (rows of `pieces`, published under `flags`), in index order.  Every consumer thread
        // polls its share of the flags (one thread polling them one after the other paid a global round trip per
        // piece) and resets what it saw — a flag has a single consumer, so the reset is safe for CUDA-graph replays;
        // the pieces are then fetched eight at a time: the loads of a batch are independent, so the cost is a round
        // trip per batch, not per piece.
        auto collect = [&](float (&sum)[kSlabs][kCols], const float *pieces, int *flags, int64_t lo, int64_t hi,
                           bool skip_rowless) {
            if (lo > hi) return;
            for (int64_t m = lo + tid; m <= hi; m += n_consumers) {
                if (skip_rowless && cta_row_begin(m, n_rows, G) == cta_row_begin(m + 1, n_rows, G)) continue;
                while (ld_acquire(flags + m) == 0) {}
                flags[m] = 0;
            }
            consumer_barrier(n_consumers);
            constexpr int kBatch = 8;
#pragma unroll
            for (int j = 0; j < kSlabs; ++j) {
                const int slab = tid + j * n_consumers;
                if (slab >= p.slabs_per_row) continue;
                const float *piece = pieces + (size_t)slab * kCols;
                for (int64_t m0 = lo; m0 <= hi; m0 += kBatch) {
                    float v[kBatch][kCols];
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        const int64_t m = m0 + b;
                        const bool live = m <= hi && !(skip_rowless && cta_row_begin(m, n_rows, G) == cta_row_begin(m + 1, n_rows, G));
#pragma unroll
                        for (int k = 0; k < kCols; k += 4) {
                            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (live) x = __ldcg(reinterpret_cast<const float4 *>(piece + (size_t)m * p.dim + k));
                            v[b][k] = x.x, v[b][k + 1] = x.y, v[b][k + 2] = x.z, v[b][k + 3] = x.w;
                        }
                    }
#pragma unroll
                    for (int b = 0; b < kBatch; ++b)
#pragma unroll
                        for (int k = 0; k < kCols; ++k) sum[j][k] += v[b][k]; // + 0 for the slots past hi
                }
            }
        };
        // publish sum[j][k] as piece `row` of the scratch (release flag by one thread after the barrier: the release is
        // cumulative, so it also covers the other consumers' stores that the barrier ordered before it)
        auto publish = [&](const float (&sum)[kSlabs][kCols], int64_t row_idx) {
            float *dst = p.head + (size_t)row_idx * p.dim;
#pragma unroll
            for (int j = 0; j < kSlabs; ++j) {
                const int slab = tid + j * n_consumers;
                if (slab < p.slabs_per_row)
#pragma unroll
                    for (int k = 0; k < kCols; k += 4)
                        *reinterpret_cast<float4 *>(dst + (size_t)slab * kCols + k) =
                            make_float4(sum[j][k], sum[j][k + 1], sum[j][k + 2], sum[j][k + 3]);
            }
            consumer_barrier(n_consumers);
            if (tid == 0) st_release(p.head_flag + row_idx, 1);
        };

        // flush the accumulators for the segment that ends (or is cut) at `row_end`.
        // Ownership rule for a segment cut by CTA boundaries: the CTA that holds its FIRST row owns it.
        // Every other CTA publishes its piece as soon as it has it and never waits before publishing (a group leader
        // waits, after its own last row, for CTAs with a higher index only); the owner waits only at the very end of
        // its own rows.  So no CTA ever waits on a waiter.
        auto flush = [&](int64_t row_end) {
            if (!in_gap) {
                const bool starts_before = seg_begin < r0;
                const bool ends_after = seg_end > row_end; // only possible when row_end == r1
                const float nrows = (float)(seg_end - seg_begin);
                if (starts_before) {
                    // end piece (or, when it also ends_after, a middle piece) of an earlier CTA's segment
                    float sum[kSlabs][kCols];
#pragma unroll
                    for (int j = 0; j < kSlabs; ++j)
#pragma unroll
                        for (int k = 0; k < kCols; ++k) sum[j][k] = reduce_acc(j, k);
                    const bool leader = ends_after && every_cta_has_rows && (c % kGroup) == 0 &&
                                        last_cta_of(seg_end) >= c + kGroup - 1;
                    if (leader) { // this CTA's group lies wholly inside the segment: one group piece instead of 16
                        collect(sum, p.head, p.head_flag, c + 1, c + kGroup - 1, false);
                        publish(sum, p.group_base + c / kGroup);
                    } else {
                        publish(sum, c);
                    }
                } else if (!ends_after) {
#pragma unroll
                    for (int j = 0; j < kSlabs; ++j) {
                        float sum[kCols];
#pragma unroll
                        for (int k = 0; k < kCols; ++k) sum[k] = reduce_acc(j, k);
                        write_pooled(j, sum, nrows);
                    }
                } else {
                    // this CTA owns a segment that continues into later CTAs: add their pieces in CTA order
                    const int64_t ce = last_cta_of(seg_end);
                    // the threads polled the flag at slightly different times: take the early path only if all saw it
                    const bool early = kEarlyCarry && ce == c + 1 && consumer_barrier_and(n_consumers, carry_flag != 0);
                    if (early) {
                        if (tid == 0) p.head_flag[c + 1] = 0; // single consumer resets: safe for CUDA-graph replays
                        float sum[kCols];
#pragma unroll
                        for (int k = 0; k < kCols; ++k) sum[k] = reduce_acc(0, k) + carry[k];
                        write_pooled(0, sum, nrows);
                    } else {
                        float sum[kSlabs][kCols];
#pragma unroll
                        for (int j = 0; j < kSlabs; ++j)
#pragma unroll
                            for (int k = 0; k < kCols; ++k) sum[j][k] = reduce_acc(j, k);
                        const int64_t g0 = c / kGroup + 1;                 // first aligned group behind this CTA
                        const int64_t g1 = (ce + 1) / kGroup - 1;          // last group that ends at or before ce
                        if (every_cta_has_rows && g0 <= g1) {
                            collect(sum, p.head, p.head_flag, c + 1, g0 * kGroup - 1, false);
                            collect(sum, p.head + (size_t)p.group_base * p.dim, p.head_flag + p.group_base, g0, g1, false);
                            collect(sum, p.head, p.head_flag, (g1 + 1) * kGroup, ce, false);
                        } else {
                            collect(sum, p.head, p.head_flag, c + 1, ce, true);
                        }
#pragma unroll
                        for (int j = 0; j < kSlabs; ++j) write_pooled(j, sum[j], nrows);
                    }
                }
            }
            clear_acc();
        };

        int64_t row = r0;
        const int64_t ch_carry = n_chunks >= 2 ? n_chunks - 2 : 0;
        for (int64_t ch = 0; ch < n_chunks; ++ch) {
            const int s = (int)(ch % kStages);
            if (ch == ch_carry) prefetch_carry();
            mbar_wait(&s_full[s], (uint32_t)((ch / kStages) & 1));
#ifdef AAT_POOL_TRACE
            if (ch == 0) POOL_TRACE(2); // first stage landed
#endif
            const unsigned char *stage = smem_raw + s * stage_stride;
            const int64_t chunk_end = (row + p.rows_per_stage < r1) ? row + p.rows_per_stage : r1;
            int rr = 0; // row within the stage
            while (row < chunk_end) {
                while (row >= seg_end) { // crossed a boundary: finish the segment, move to the next non-empty one
                    flush(row);
                    do {
                        ++seg;
                        load_segment();
                    } while (!in_gap && seg_end == seg_begin);
                }
                const int64_t stop = seg_end < chunk_end ? seg_end : chunk_end;
                const int run = (int)(stop - row);
                if (!in_gap) {
                    const unsigned char *base = stage + (size_t)rr * p.row_bytes + (size_t)tid * 16;
                    int i = 0;
                    for (; i + 4 <= run; i += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int j = 0; j < kSlabs; ++j) {
                                if (tid + j * n_consumers < p.slabs_per_row) {
                                    float v[kCols];
                                    S::load(base + (size_t)(i + u) * p.row_bytes + (size_t)j * n_consumers * 16, v);
#pragma unroll
                                    for (int k = 0; k < kCols; ++k) acc[j][u][k] += v[k];
                                }
                            }
                    }
                    for (; i < run; ++i)
#pragma unroll
                        for (int j = 0; j < kSlabs; ++j) {
                            if (tid + j * n_consumers < p.slabs_per_row) {
                                float v[kCols];
                                S::load(base + (size_t)i * p.row_bytes + (size_t)j * n_consumers * 16, v);
#pragma unroll
                                for (int k = 0; k < kCols; ++k) acc[j][0][k] += v[k];
                            }
                        }
                }
                row += run;
                rr += run;
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&s_empty[s]);
        }
        POOL_TRACE(3); // last stage consumed
        flush(r1); // the segment still open at the end of this CTA's rows
        POOL_TRACE(4); // carry collected
    }

    if (kColsum) {
#pragma unroll
        for (int j = 0; j < kSlabs; ++j) {
            const int slab = tid + j * n_consumers;
            if (slab < p.slabs_per_row)
#pragma unroll
                for (int k = 0; k < kCols; ++k) p.colsum[(size_t)c * p.dim + (size_t)slab * kCols + k] = csum[kColsum ? j : 0][k];
        }
    }
#ifdef AAT_POOL_TRACE
    POOL_TRACE(5); // exit
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_pool_trace[blockIdx.x * 8 + 6] = smid;
    }
#endif
}

// colsum_out[d] (+)= sum over the pool CTAs of their per-CTA column sums; colsum_out[dim] (+)= S.
// Block = 32 columns x 32 slices of the CTA axis (coalesced 256-byte rows).  Every thread issues all of its
// (<= 16) loads before it adds anything, so the kernel costs one memory latency, not one per row; the slice
// partials are then added in slice order, so the result does not depend on scheduling.
constexpr int kReduceSlices = 32;
constexpr int kReducePer = 16; // rows per thread: supports up to 512 pool CTAs
__global__ void __launch_bounds__(32 * kReduceSlices)
colsum_reduce_kernel(const double *partial, int n_ctas, int dim, int64_t n_seg, const int64_t *n_seg_dev, double *out,
                     bool accumulate)
{
    __shared__ double s_part[kReduceSlices][33];
    pdl_wait(); // the pool kernel's per-CTA sums
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + lane;
    const int per = (n_ctas + kReduceSlices - 1) / kReduceSlices;
    const int c0 = slice * per;
    double v[kReducePer];
#pragma unroll
    for (int k = 0; k < kReducePer; ++k) {
        const int c = c0 + k;
        v[k] = (k < per && c < n_ctas && d < dim) ? partial[(size_t)c * dim + d] : 0.0;
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kReducePer; ++k) s += v[k];
    s_part[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && d < dim) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kReduceSlices; ++k) t += s_part[k][lane];
        out[d] = accumulate ? out[d] + t : t;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double count = (double)(n_seg_dev ? min(*n_seg_dev, n_seg) : n_seg);
        out[dim] = accumulate ? out[dim] + count : count;
    }
}

__global__ void colsum_accumulate_kernel(double *acc, const double *colsum, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += colsum[i];
}

__global__ void colsum_finalize_kernel(const double *acc, int dim, float *mean)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) mean[i] = (float)(acc[i] / acc[dim]);
}

template <typename EmbT, int kSlabs>
int launch_typed(aat_ctx *ctx, const PoolScratch &ps, PoolParams &p, size_t smem, bool colsum, int ctas_per_sm,
                 cudaStream_t stream, int *grid_out)
{
    const int threads = p.n_consumers + 32;
    auto kernel = colsum ? pool_kernel<EmbT, kSlabs, true> : pool_kernel<EmbT, kSlabs, false>;
    // persistent grid, sized from the occupancy the driver reports for this very instantiation
    int per_sm = 0;
    AAT_CUDA_CHECK(prepare_kernel(ctx, kernel, threads, smem, &per_sm));
    AAT_REQUIRE(per_sm >= 1, AAT_ERR_UNSUPPORTED, "aat_segment_mean_pool: kernel does not fit on an SM");
    if (per_sm > ctas_per_sm) per_sm = ctas_per_sm;
    int grid = ctx->num_sms * per_sm;
    if (grid > ps.max_ctas) grid = ps.max_ctas;
    // (More, smaller CTAs handed out in waves by the hardware, as a cheap form of dynamic balancing, were tried:
    //  2 / 3 waves cost 30.2 -> 34.4 / 43.5 us at config-2 size, every new CTA pays the ramp again, and gain 0.4 % at
    //  config-4 size; gpurun r2, AAT_POOL_WAVES experiment.)
    // small inputs: give every CTA at least two stages of rows, otherwise one segment spans dozens of CTAs
    // and its owner spends longer collecting pieces than streaming
    const int64_t min_rows = 2 * (int64_t)p.rows_per_stage;
    const int64_t by_rows = (p.n_rows + min_rows - 1) / min_rows;
    if (by_rows < grid) grid = by_rows < 1 ? 1 : (int)by_rows;
#ifdef AAT_EXPERIMENTS
    if (const char *e = getenv("AAT_POOL_GRID")) // profiles/pool_grid.py: what a few SMs alone can stream
        if (atoi(e) > 0 && atoi(e) < grid) grid = atoi(e);
#endif
    *grid_out = grid;
    ProfileScope prof(ctx, AAT_K_POOL, stream); // the streaming kernel alone (not the colsum reduce)
    AAT_CUDA_CHECK(launch_pdl(kernel, dim3(grid), dim3(threads), smem, stream, p));
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

template <typename EmbT>
int launch_slabs(aat_ctx *ctx, const PoolScratch &ps, PoolParams &p, int slabs, size_t smem, bool colsum, int ctas_per_sm,
                 cudaStream_t stream, int *grid_out)
{
    switch (slabs) {
    case 1: return launch_typed<EmbT, 1>(ctx, ps, p, smem, colsum, ctas_per_sm, stream, grid_out);
    case 2: return launch_typed<EmbT, 2>(ctx, ps, p, smem, colsum, ctas_per_sm, stream, grid_out);
    default: return launch_typed<EmbT, 4>(ctx, ps, p, smem, colsum, ctas_per_sm, stream, grid_out);
    }
}

} // namespace

int pool_scratch_init(int num_sms, PoolScratch *ps)
{
    ps->max_ctas = num_sms * kMaxCtasPerSm;
    if (ps->max_ctas > 512) ps->max_ctas = 512; // colsum_reduce_kernel covers 32 slices x 16 rows
    ps->max_dim = 4096;
    const size_t pieces = (size_t)ps->max_ctas + (size_t)ps->max_ctas / kGroup + 1; // per-CTA pieces, then group pieces
    AAT_CUDA_CHECK(cudaMalloc(&ps->head, sizeof(float) * pieces * ps->max_dim));
    AAT_CUDA_CHECK(cudaMalloc(&ps->head_flag, sizeof(int) * pieces));
    AAT_CUDA_CHECK(cudaMalloc(&ps->colsum, sizeof(double) * (size_t)ps->max_ctas * ps->max_dim));
    AAT_CUDA_CHECK(cudaMemset(ps->head_flag, 0, sizeof(int) * pieces));
    return AAT_OK;
}

void pool_scratch_free(PoolScratch *ps)
{
    cudaFree(ps->head);
    cudaFree(ps->head_flag);
    cudaFree(ps->colsum);
    *ps = PoolScratch{};
}

static int launch_mean_pool_on(aat_ctx *ctx, const PoolScratch &ps, const void *emb, int emb_dtype, int64_t n_rows,
                               int32_t dim, const int64_t *seg_off, int64_t n_seg, const int64_t *n_seg_dev, float *out,
                               double *colsum, int flags, cudaStream_t stream)
{
    int esize;
    switch (emb_dtype) {
    case AAT_F32: esize = 4; break;
    case AAT_F16:
    case AAT_BF16: esize = 2; break;
    default:
        AAT_REQUIRE(false, AAT_ERR_UNSUPPORTED, "aat_segment_mean_pool: embedding dtype must be F32, F16 or BF16");
    }
    const bool colsum_accumulate = (flags & AAT_POOL_ACCUMULATE) != 0;
    AAT_REQUIRE(dim > 0 && n_rows >= 0 && n_seg >= 0, AAT_ERR_INVALID, "aat_segment_mean_pool: negative size");
    AAT_REQUIRE(!(flags & AAT_POOL_ROWS_FROM_DEVICE) || n_seg_dev != nullptr, AAT_ERR_INVALID,
                "aat_segment_mean_pool: AAT_POOL_ROWS_FROM_DEVICE needs n_seg_dev");
    const int64_t row_bytes = (int64_t)dim * esize;
    AAT_REQUIRE(row_bytes % 16 == 0, AAT_ERR_UNSUPPORTED,
                "aat_segment_mean_pool: dim * sizeof(element) = %lld must be a multiple of 16", (long long)row_bytes);
    AAT_REQUIRE(row_bytes <= 16 * kMaxConsumers * kMaxSlabs && dim <= ps.max_dim, AAT_ERR_UNSUPPORTED,
                "aat_segment_mean_pool: dim %d too large (row must be <= %d bytes, dim <= %d)", dim,
                16 * kMaxConsumers * kMaxSlabs, ps.max_dim);
    AAT_REQUIRE((reinterpret_cast<uintptr_t>(emb) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                AAT_ERR_INVALID, "aat_segment_mean_pool: emb_dev and out_dev must be 16-byte aligned");
    if (n_seg == 0) {
        if (colsum && !colsum_accumulate)
            AAT_CUDA_CHECK(cudaMemsetAsync(colsum, 0, sizeof(double) * (size_t)(dim + 1), stream));
        return AAT_OK;
    }
    AAT_REQUIRE(emb != nullptr || n_rows == 0, AAT_ERR_INVALID, "aat_segment_mean_pool: emb_dev is NULL");
    AAT_REQUIRE(seg_off != nullptr && out != nullptr, AAT_ERR_INVALID, "aat_segment_mean_pool: NULL seg_off/out");

    PoolParams p{};
    p.emb = static_cast<const unsigned char *>(emb);
    p.seg_off = seg_off;
    p.n_seg_dev = n_seg_dev;
    p.out = out;
    p.head = ps.head;
    p.head_flag = ps.head_flag;
    p.colsum = ps.colsum;
    p.group_base = ps.max_ctas;
    p.n_rows = n_rows;
    p.n_seg = n_seg;
    p.dim = dim;
    p.row_bytes = (int)row_bytes;
    p.slabs_per_row = (int)(row_bytes / 16);
    p.rows_from_dev = (flags & AAT_POOL_ROWS_FROM_DEVICE) ? 1 : 0;
    // the row split is only known after the dependency wait when the row count lives on the device
    p.pre_stages = ((flags & AAT_POOL_EMB_READY) && !p.rows_from_dev) ? kPreStages : 0;
    int consumers = ((p.slabs_per_row + 31) / 32) * 32;
    int slabs = 1;
    while (consumers > kMaxConsumers) {
        slabs *= 2;
        consumers = (((p.slabs_per_row + slabs - 1) / slabs + 31) / 32) * 32;
    }
    p.n_consumers = consumers;
    p.rows_per_stage = (int)(kStageBytes / row_bytes);
    if (p.rows_per_stage < 1) p.rows_per_stage = 1;
    const size_t smem = (size_t)kStages * p.rows_per_stage * row_bytes;

    int rc, grid = 0;
    const bool want_colsum = colsum != nullptr;
    // AAT_POOL_SHARE_SMS: half the persistent grid, so that two launches (or a launch and a log-mel CTA) share an SM
    const int ctas_per_sm = (flags & AAT_POOL_SHARE_SMS) ? 1 : kMaxCtasPerSm;
    if (emb_dtype == AAT_F32)
        rc = launch_slabs<float>(ctx, ps, p, slabs, smem, want_colsum, ctas_per_sm, stream, &grid);
    else if (emb_dtype == AAT_F16)
        rc = launch_slabs<__half>(ctx, ps, p, slabs, smem, want_colsum, ctas_per_sm, stream, &grid);
    else
        rc = launch_slabs<__nv_bfloat16>(ctx, ps, p, slabs, smem, want_colsum, ctas_per_sm, stream, &grid);
    if (rc != AAT_OK) return rc;
    if (want_colsum) {
        AAT_MAX_SMEM_CARVEOUT(colsum_reduce_kernel);
        AAT_CUDA_CHECK(launch_pdl(colsum_reduce_kernel, dim3((dim + 31) / 32), dim3(32 * kReduceSlices), 0, stream,
                                  (const double *)ps.colsum, grid, (int)dim, n_seg, n_seg_dev, colsum,
                                  colsum_accumulate));
        AAT_LAUNCH_CHECK();
    }
    return AAT_OK;
}

int launch_mean_pool(aat_ctx *ctx, const aat_plan *plan, const void *emb, int emb_dtype, int64_t n_rows, int32_t dim,
                     const int64_t *seg_off, int64_t n_seg, const int64_t *n_seg_dev, float *out, double *colsum,
                     int flags, cudaStream_t stream)
{
    if (plan != nullptr && plan->pool.head != nullptr) // the plan's own scratch: one launch per plan in flight
        return launch_mean_pool_on(ctx, plan->pool, emb, emb_dtype, n_rows, dim, seg_off, n_seg, n_seg_dev, out, colsum,
                                   flags, stream);
    // No plan: the context's scratch.  Launches that share it are ordered against each other whatever streams they
    // are on: each waits for the event recorded behind the previous one.  (Inside a stream capture the event cannot
    // be used — a captured stream may not depend on uncaptured work — so captured plan-less launches must not run
    // concurrently with other plan-less launches; name a plan there.)
    std::lock_guard<std::mutex> lock(ctx->pool_mutex);
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    AAT_CUDA_CHECK(cudaStreamIsCapturing(stream, &capturing));
    const bool use_event = capturing == cudaStreamCaptureStatusNone;
    if (use_event && ctx->pool_done_recorded) AAT_CUDA_CHECK(cudaStreamWaitEvent(stream, ctx->pool_done, 0));
    const int rc = launch_mean_pool_on(ctx, ctx->pool, emb, emb_dtype, n_rows, dim, seg_off, n_seg, n_seg_dev, out,
                                       colsum, flags, stream);
    if (rc != AAT_OK) return rc;
    if (use_event) {
        AAT_CUDA_CHECK(cudaEventRecord(ctx->pool_done, stream));
        ctx->pool_done_recorded = true;
    }
    return AAT_OK;
}

#ifdef AAT_POOL_TRACE
extern "C" __attribute__((visibility("default"))) int aat_debug_pool_trace(unsigned long long *out_host, int n_ctas)
{
    return (int)cudaMemcpyFromSymbol(out_host, g_pool_trace, sizeof(unsigned long long) * 8 * (size_t)n_ctas);
}
#endif

int launch_colsum_accumulate(double *acc, const double *colsum, int32_t dim, cudaStream_t stream)
{
    colsum_accumulate_kernel<<<(dim + 1 + 127) / 128, 128, 0, stream>>>(acc, colsum, dim + 1);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

int launch_colsum_finalize(const double *acc, int32_t dim, float *mean, cudaStream_t stream)
{
    colsum_finalize_kernel<<<(dim + 127) / 128, 128, 0, stream>>>(acc, dim, mean);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat

AAT_TIMELINE_EXPORT(pool, aat::)
