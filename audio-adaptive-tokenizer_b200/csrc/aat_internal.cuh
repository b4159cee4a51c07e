// Internal definitions shared by the kernels and the C ABI (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "aat_b200.h"

namespace aat {

constexpr int kNfft = 400;          // the FFT kernel is specialised for 400 = 20 x 20
constexpr int kBins = kNfft / 2 + 1; // 201
constexpr int kMaxMels = 128;
constexpr int kMaxRunningMean = 2040; // boundary kernel: cumsum ring (8192) >= running_mean_points + 2 + 3 chunks

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launch_count;

#define AAT_CUDA_CHECK(expr)                                                                         \
    do {                                                                                             \
        cudaError_t err__ = (expr);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            aat::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return AAT_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

#define AAT_REQUIRE(cond, status, ...)  \
    do {                                \
        if (!(cond)) {                  \
            aat::set_error(__VA_ARGS__); \
            return (status);            \
        }                               \
    } while (0)

// Every kernel of the library asks for the same (maximum) shared-memory carve-out, so the SMs are not
// drained and reconfigured between the kernels of a step (they differ widely in shared-memory footprint).
// Function attributes and occupancy are per-device facts that do not change between launches: they are set / asked
// once per (context, kernel) and remembered — ten driver calls per step were half of the host time of a step.
struct KernelSetup {
    size_t max_smem = 0;   // cudaFuncAttributeMaxDynamicSharedMemorySize already granted
    bool carveout = false; // cudaFuncAttributePreferredSharedMemoryCarveout already set
    int occ_threads = 0;   // occupancy answer for (occ_threads, occ_smem)
    size_t occ_smem = 0;
    int occ = 0;
};
#define AAT_MAX_SMEM_CARVEOUT(kernel) AAT_CUDA_CHECK(aat::prepare_kernel(ctx, kernel, 0, 0, nullptr))

#define AAT_LAUNCH_CHECK()                         \
    do {                                           \
        aat::g_launch_count.fetch_add(1);          \
        AAT_CUDA_CHECK(cudaGetLastError());        \
    } while (0)

// The dense (bins, mels) float64 filter bank in the banded form the log-mel kernel walks.  Every filter covers a
// run of consecutive bins.  In the kernel thread (frame f, group q) evaluates the filters q, q + kMelGroups, ... of
// its frame.  The taps of a group are ONE stream of chunks of four (weights zero-padded to whole chunks), so the kernel
// runs a single rolled loop per thread with the next chunk's operands in flight, instead of dispatching on every
// filter's length.  A chunk's descriptor says where the NEXT chunk's power values start and whether this chunk ends a
// filter.  The two groups that share a warp (q even and q + 1) own neighbouring filters, which are padded to a common
// chunk count so that the end-of-filter branch is warp-uniform.
constexpr int kMelGroups = 10;    // thread groups of the mel phase (160 threads / 16 frames)
constexpr int kMelChunk = 4;      // taps per chunk = independent FMA chains per filter
constexpr int kMelDescHeader = 32; // uint16 header: per group q [3q] first chunk, [3q+1] chunks, [3q+2] first bin
constexpr uint16_t kMelDescLast = 0x8000; // descriptor bit: the chunk is the last one of its filter (bits 0-7: next bin)
constexpr int kMelMaxWeights = 1536; // padded weights that fit the kernel's shared-memory budget at two CTAs per SM
struct MelSchedule {
    int n_mels = 0;
    int nnz = 0;        // total band length of the filters
    int n_weights = 0;  // padded weights (doubles): kMelChunk per chunk, one chunk of padding at the end
    int n_desc = 0;     // uint16 entries of chunk_desc: header + one per chunk + 2 of padding
    uint16_t *chunk_desc = nullptr; // device [n_desc]
    double *weight = nullptr;       // device [n_weights], group-major chunk streams
};

// One tile (kMelFramesPerTile consecutive frames of one utterance) of the log-mel kernel.
struct alignas(16) MelTile {
    int64_t src;      // element index (in the packed waveform) of the first staged sample; < wave_off for the first tile
    int64_t n;        // samples of the utterance
    int64_t wave_off; // first sample of the utterance
    int64_t mel_off;  // element index of mel[0][first frame of the tile] in the packed log-mel
    int64_t amp_off;  // index of the tile's first frame in the per-frame arrays
    int32_t T;        // frames of the utterance (row stride of its mel block)
    int32_t valid;    // frames of this tile that exist (1..kMelFramesPerTile); 0 marks "no tile"
    int32_t interior; // the whole staged range lies inside the utterance (no reflection needed)
    int32_t utt;      // utterance index (per-utterance statistics of the fused z-score)
};
static_assert(sizeof(MelTile) == 64, "MelTile is copied as four 16-byte pieces");

// Scratch for the pool kernel's cross-CTA partial sums.  One launch at a time may use a scratch block: a plan owns
// one (a plan has one launch in flight at a time anyway), and the context owns the block used by calls without a plan,
// whose launches are ordered against each other with an event (see launch_mean_pool).
struct PoolScratch {
    int max_ctas = 0;
    int max_dim = 0;
    float *head = nullptr;      // [max_ctas, max_dim]
    int *head_flag = nullptr;   // [max_ctas]
    double *colsum = nullptr;   // [max_ctas, max_dim + 1]
};

// Event-pair ring for aat_profile_* (created lazily on enable).
struct Profiler {
    uint32_t mask = 0;
    std::vector<cudaEvent_t> start, stop;
    std::vector<int> kernel;
    size_t used = 0;
    int every = 1;                 // record every n-th launch of a kernel
    int64_t seen[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // launches of each kernel id since the last enable
};

} // namespace aat

struct aat_ctx {
    int device = 0;
    int num_sms = 0;
    aat_config cfg{};
    double *window_half = nullptr; // device [400], 0.5 * window (exact scaling, folds the /2 of the two-frame split)
    double2 *twiddle = nullptr;    // device [7 * 20], W_400^(k1 * n2) for k1 = 1, 2, 3, 4, 5, 10, 15 (row-major over n2)
    double2 *log_table = nullptr;  // device [32], (1/c_i, -log10(1/c_i)) for the log-mel kernel's log10
    aat::MelSchedule mel{};
    aat::PoolScratch pool{};         // scratch of pool launches that name no plan
    std::mutex pool_mutex;           // ... which are ordered against each other: each waits for `pool_done`, recorded
    cudaEvent_t pool_done = nullptr; //     behind the previous one (on whatever stream that was)
    bool pool_done_recorded = false;
    // staging for aat_host_* entry points (grown on demand, never inside stream capture)
    void *dev_scratch = nullptr;
    size_t dev_scratch_bytes = 0;
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaStream_t host_stream = nullptr;
    // aat_host_* entry points: serialised (they share the staging buffers; ctypes callers release the GIL) and
    // backed by a small most-recently-used cache of single-utterance plans keyed by the sample count
    std::mutex host_mutex;
    std::vector<std::pair<int64_t, aat_plan *>> host_plans;
    aat::Profiler prof{};
    std::mutex setup_mutex;
    std::map<const void *, aat::KernelSetup> kernel_setup; // keyed by the kernel's host function pointer
};

struct aat_plan {
    aat_ctx *ctx = nullptr;
    int32_t n_utts = 0;
    int64_t total_samples = 0, total_frames = 0, total_seg_slots = 0;
    int64_t max_frames = 0;
    int32_t mel_tiles = 0; // CTAs of the log-mel kernel
    std::vector<int64_t> h_n_samples, h_wave_off, h_frame_off, h_seg_slot_off;
    // device tables
    int64_t *d_n_samples = nullptr;     // [B]
    int64_t *d_wave_off = nullptr;      // [B+1]
    int64_t *d_frame_off = nullptr;     // [B+1]
    int64_t *d_seg_slot_off = nullptr;  // [B+1]
    aat::MelTile *d_mel_tile = nullptr; // [mel_tiles] tile descriptors of the log-mel kernel
    int32_t *d_mel_sched = nullptr;     // [4] self-resetting counters: the log-mel kernel's tile and exit counters, the
                                        // boundary kernel's completion ticket (one launch per plan in flight at a time)
    // scratch written by the boundaries kernel for its fused frame-CSR epilogue
    int64_t *d_seg_local = nullptr;     // [total_seg_slots]
    int64_t *d_utt_frames = nullptr;    // [B] look-back words of the boundary kernel's CSR epilogue (zero between launches)
    // waveform normalisation (aat_normalize): 4096-sample chunks
    int32_t norm_chunks = 0;
    int32_t *d_chunk_utt = nullptr;     // [norm_chunks]
    int32_t *d_chunk_first = nullptr;   // [B+1]
    double *d_norm_partial = nullptr;   // [norm_chunks, 3] (n, mean, M2)
    double *d_norm_stats = nullptr;     // [B, 2] (mean, population variance)
    aat::PoolScratch pool{};            // cross-CTA scratch of the pool kernel (empty for the cached host-API plans)
    // on-device synthetic inputs (aat_synth_waveforms): slots of the per-utterance burst tables in the caller's workspace
    int64_t total_bursts = 0;
    int64_t *d_burst_off = nullptr;     // [B+1]
};

namespace aat {

// kernels' host launchers (defined in the .cu files)
int launch_logmel(aat_ctx *ctx, const aat_plan *plan, const void *wave, int wave_dtype, const double *znorm_stats,
                  float *mel, float *amp, cudaStream_t stream);
int launch_boundaries(aat_ctx *ctx, const aat_plan *plan, const float *mel, const float *amp, int64_t *seg_start,
                      int64_t *seg_len, int32_t *seg_count, int64_t *minima, int32_t *minima_count, int32_t *status,
                      int64_t *seg_off, int64_t *n_seg, int64_t *utt_seg_off, cudaStream_t stream);
int launch_amplitude(aat_ctx *ctx, const aat_plan *plan, const float *mel, float *amp, cudaStream_t stream);
int launch_process_boarders(aat_ctx *ctx, int64_t n_samples, const int64_t *boarders, int64_t n_boarders,
                            int64_t *seg_start, int64_t *seg_len, int64_t capacity, int32_t *seg_count,
                            int32_t *status, cudaStream_t stream);
int launch_segment_frame_csr(aat_ctx *ctx, const aat_plan *plan, const int64_t *seg_len, const int32_t *seg_count,
                             int64_t *seg_off, int64_t *n_seg, int64_t *utt_seg_off, cudaStream_t stream);
int launch_utterance_frame_csr(const aat_plan *plan, const int64_t *seg_start, const int32_t *seg_count,
                               const int64_t *utt_seg_off, int64_t *seg_off, int64_t *n_seg, cudaStream_t stream);
int launch_mean_pool(aat_ctx *ctx, const aat_plan *plan, const void *emb, int emb_dtype, int64_t n_rows, int32_t dim,
                     const int64_t *seg_off, int64_t n_seg, const int64_t *n_seg_dev, float *out, double *colsum,
                     int flags, cudaStream_t stream);
int launch_colsum_accumulate(double *acc, const double *colsum, int32_t dim, cudaStream_t stream);
int launch_colsum_finalize(const double *acc, int32_t dim, float *mean, cudaStream_t stream);
int launch_normalize(aat_ctx *ctx, const aat_plan *plan, const void *wave, int in_dtype, int mode, void *out,
                     int out_dtype, double *stats, cudaStream_t stream);
int launch_pad_boarders(const aat_plan *plan, const int64_t *seg_len, const int32_t *seg_count, int64_t s_max,
                        int64_t *boarders, int64_t *mask, int32_t *status, cudaStream_t stream);
int launch_scatter_segments(const float *wave, int64_t n_max, int32_t n_utts, const int64_t *boarders, int64_t s_max,
                            int64_t max_frames, float *out, float *mask, int32_t *status, cudaStream_t stream);
int launch_scatter_mel_segments(aat_ctx *ctx, const aat_plan *plan, int32_t n_utts, const float *mel,
                                const int64_t *mel_elem_off, const int64_t *mel_frames, const int64_t *mel_row_stride,
                                const int64_t *boarders, int64_t s_max, int64_t max_items, float *out, int32_t *status,
                                cudaStream_t stream);
int launch_normalize_padded(aat_ctx *ctx, const aat_plan *plan, const void *wave, int in_dtype, int mode, float *out,
                            int64_t n_max, int32_t *mask, double *stats, cudaStream_t stream);
int launch_masked_mean_pool(const void *emb, int emb_dtype, int64_t n_rows, int64_t seq_len, int32_t dim,
                            const int64_t *mask, float *out, int64_t *row_mask, cudaStream_t stream);
int64_t synth_burst_capacity(int sampling_rate, int64_t n_samples);
size_t synth_workspace_bytes(const aat_plan *plan);
int launch_synth_waveforms(aat_ctx *ctx, const aat_plan *plan, uint64_t seed_base, int64_t utt_index_base, float *wave,
                           void *workspace, cudaStream_t stream);
int launch_synth_normal(aat_ctx *ctx, float *out, int64_t n, uint64_t seed, cudaStream_t stream);
constexpr int kNormChunk = 4096;
int logmel_tables_init(aat_ctx *ctx);
int pool_scratch_init(int num_sms, PoolScratch *ps);
void pool_scratch_free(PoolScratch *ps);

constexpr int kMelFramesPerTile = 16; // frames one CTA of the log-mel kernel produces

// -DAAT_TIMELINE (the timeline build of profiles/step_timeline.py, never the product library): thread 0 of every CTA of
// the three path kernels records {start, end, SM << 32 | CTA} of its life (global timer, ns) in a ring of its file.
#ifdef AAT_TIMELINE
constexpr unsigned kTimelineRing = 1u << 16;
struct TimelineScope {
    unsigned long long t0;
    unsigned long long *ring;
    unsigned *count;
    __device__ static unsigned long long now()
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    }
    __device__ TimelineScope(unsigned long long *r, unsigned *c) : t0(now()), ring(r), count(c) {}
    __device__ ~TimelineScope()
    {
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const unsigned i = atomicAdd(count, 1u) & (kTimelineRing - 1);
            ring[3 * i] = t0, ring[3 * i + 1] = now(), ring[3 * i + 2] = ((unsigned long long)smid << 32) | blockIdx.x;
        }
    }
};
#define AAT_TIMELINE_STORAGE(name)                                     \
    __device__ unsigned long long g_tl_ring_##name[3 * kTimelineRing]; \
    __device__ unsigned g_tl_count_##name;
// at the end of the file, outside every namespace (NS = the namespaces the storage sits in)
#define AAT_TIMELINE_EXPORT(name, NS)                                                                                         \
    extern "C" __attribute__((visibility("default"))) int aat_debug_timeline_##name(unsigned long long *out_host,            \
                                                                                    unsigned *count_host)                    \
    {                                                                                                                         \
        if (cudaMemcpyFromSymbol(count_host, NS g_tl_count_##name, sizeof(unsigned)) != cudaSuccess) return -1;               \
        return (int)cudaMemcpyFromSymbol(out_host, NS g_tl_ring_##name, sizeof(unsigned long long) * 3 * aat::kTimelineRing); \
    }
#define AAT_TIMELINE_SCOPE(name) TimelineScope tl_scope__(g_tl_ring_##name, &g_tl_count_##name)
#else
#define AAT_TIMELINE_STORAGE(name)
#define AAT_TIMELINE_EXPORT(name, NS)
#define AAT_TIMELINE_SCOPE(name) \
    do {                         \
    } while (0)
#endif


// Programmatic dependent launch (PDL) between the kernels of a step.  A kernel launched through launch_pdl may
// start while its predecessor in the stream is still running; it must call pdl_wait() before it touches
// anything the predecessor writes (and before it writes anything the predecessor reads), and calls
// pdl_launch_dependents() afterwards so that its own successor may be scheduled early in turn.  Because
// every kernel waits before it triggers, completion is transitive along the chain.  Both instructions are
// no-ops when the kernel was launched normally, and a normal launch after one of these kernels is fully ordered.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#ifdef __CUDACC__
// (x - mean) / (std + 1e-6) in float64 with the correctly rounded quotient numpy's division gives, without a division
// per sample: with r = RN(1 / d) taken once, q = RN(t * r) is within an ulp of t / d, the remainder t - q * d is exact
// in one FMA, and RN(q + rem * r) is the correctly rounded quotient (Markstein) — three operations on the FP64 pipe
// instead of the ~10 of a division.  Shared by aat_normalize and by the log-mel kernel's fused z-score.
struct Znorm {
    double mean, d, r;
    __device__ __forceinline__ static Znorm from_stats(const double *stats, int utt)
    {
        Znorm z;
        z.mean = __ldg(stats + 2 * utt);
        z.d = __dadd_rn(__dsqrt_rn(__ldg(stats + 2 * utt + 1)), 1e-6);
        z.r = __drcp_rn(z.d);
        return z;
    }
    __device__ __forceinline__ double operator()(double x) const
    {
        const double t = __dsub_rn(x, mean);
        const double q = __dmul_rn(t, r);
        return fma(fma(-q, d, t), r, q);
    }
};
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Grants `smem` bytes of dynamic shared memory to `kernel`, asks for the maximum carve-out, and (when per_sm is given)
// returns the resident CTAs per SM for (threads, smem) — each at most once per context and kernel.
template <typename K>
inline cudaError_t prepare_kernel(aat_ctx *ctx, K kernel, int threads, size_t smem, int *per_sm)
{
    std::lock_guard<std::mutex> lock(ctx->setup_mutex);
    KernelSetup &ks = ctx->kernel_setup[reinterpret_cast<const void *>(kernel)];
    cudaError_t err;
    if (smem > ks.max_smem) {
        if ((err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
        ks.max_smem = smem;
    }
    if (!ks.carveout) {
        if ((err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) !=
            cudaSuccess)
            return err;
        ks.carveout = true;
    }
    if (per_sm) {
        if (ks.occ_threads != threads || ks.occ_smem != smem) {
            int occ = 0;
            if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem)) != cudaSuccess) return err;
            ks.occ_threads = threads, ks.occ_smem = smem, ks.occ = occ;
        }
        *per_sm = ks.occ;
    }
    return cudaSuccess;
}

// RAII helper: records an event pair around one launch when profiling is enabled for `id`.
struct ProfileScope {
    aat_ctx *ctx;
    cudaStream_t stream;
    long slot = -1;
    ProfileScope(aat_ctx *c, int id, cudaStream_t s) : ctx(c), stream(s)
    {
        Profiler &p = c->prof;
        if (((p.mask >> id) & 1u) && (p.seen[id]++ % p.every) == 0) {
            // a launch that is being captured into a CUDA graph is not timed (a timing event recorded inside a capture
            // is a graph node, not a time stamp)
            cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(s, &capturing) != cudaSuccess || capturing != cudaStreamCaptureStatusNone) return;
            if (p.used < p.start.size()) {
                slot = (long)p.used++;
                p.kernel[slot] = id;
                cudaEventRecord(p.start[slot], stream);
            }
        }
    }
    ~ProfileScope()
    {
        if (slot >= 0) cudaEventRecord(ctx->prof.stop[slot], stream);
    }
};

} // namespace aat
