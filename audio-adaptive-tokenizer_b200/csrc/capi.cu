// C ABI of libaat_b200.so (include/aat_b200.h): context and plan management, argument checking,
// and the aat_host_* entry points that stage host buffers through pinned memory.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

#include "aat_internal.cuh"

namespace aat {

static thread_local char g_error[512] = "";
std::atomic<int64_t> g_launch_count{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            ok = false;
            return;
        }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// device entry points run on the context's device whatever the caller's current device is
#define AAT_DEVICE_GUARD(ctx)                  \
    DeviceGuard guard__((ctx)->device);         \
    AAT_REQUIRE(guard__.ok, AAT_ERR_CUDA, "cudaSetDevice(%d) failed", (ctx)->device)

template <typename T>
int upload(T **dst, const T *src, size_t n)
{
    AAT_CUDA_CHECK(cudaMalloc(dst, sizeof(T) * (n ? n : 1)));
    if (n) AAT_CUDA_CHECK(cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
    return AAT_OK;
}

int64_t seg_capacity(const aat_config &cfg, int64_t n)
{
    // every accepted boarder consumes >= min samples and emits <= len/max + 1 pieces; + N boarder + padded tail
    const int64_t by_min = cfg.min_segment_frames > 0 ? n / cfg.min_segment_frames : n / cfg.hop_length + 1;
    return by_min + n / cfg.max_segment_frames + 4;
}

// grow-only staging buffers for the aat_host_* entry points
int ensure_scratch(aat_ctx *ctx, size_t dev_bytes, size_t pinned_bytes)
{
    if (dev_bytes > ctx->dev_scratch_bytes) {
        if (ctx->dev_scratch) cudaFree(ctx->dev_scratch);
        ctx->dev_scratch = nullptr;
        ctx->dev_scratch_bytes = 0;
        const size_t want = dev_bytes + dev_bytes / 4 + 4096;
        AAT_CUDA_CHECK(cudaMalloc(&ctx->dev_scratch, want));
        ctx->dev_scratch_bytes = want;
    }
    if (pinned_bytes > ctx->pinned_bytes) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
        const size_t want = pinned_bytes + pinned_bytes / 4 + 4096;
        AAT_CUDA_CHECK(cudaMallocHost(&ctx->pinned, want));
        ctx->pinned_bytes = want;
    }
    return AAT_OK;
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// bump allocator over a scratch block
struct Arena {
    unsigned char *base;
    size_t off = 0;
    explicit Arena(void *b) : base(static_cast<unsigned char *>(b)) {}
    template <typename T>
    T *take(size_t n)
    {
        T *p = reinterpret_cast<T *>(base + off);
        off += align256(sizeof(T) * (n ? n : 1));
        return p;
    }
};

// dense (bins, mels) -> per-group chunk streams (see MelSchedule).  Filter m covers bins [first_m, last_m] (its first
// and last non-zero weight; interior zeros are kept so the run stays contiguous), zero-padded to whole chunks of four taps.
int build_mel_schedule(aat_ctx *ctx, const double *filters)
{
    const int M = ctx->cfg.num_mel_filters;
    std::vector<int> lo(M, -1), hi(M, -1);
    int nnz = 0;
    for (int m = 0; m < M; ++m) {
        for (int k = 0; k < kBins; ++k)
            if (filters[(size_t)k * M + m] != 0.0) {
                if (lo[m] < 0) lo[m] = k;
                hi[m] = k;
            }
        if (lo[m] >= 0) nnz += hi[m] - lo[m] + 1;
    }
    // an empty filter (band above Nyquist) is one zero tap at bin 0: 0 * NaN stays NaN as in the reference's dense product
    auto len_of = [&](int m) { return lo[m] < 0 ? 1 : hi[m] - lo[m] + 1; };
    std::vector<uint16_t> desc(kMelDescHeader, 0);
    std::vector<double> weights;
    for (int q = 0; q < kMelGroups; ++q) {
        const size_t first_chunk = weights.size() / kMelChunk;
        int prev_desc = -1; // index of the last chunk written for this group: it learns where the next chunk starts
        for (int m = q; m < M; m += kMelGroups) {
            // the filter evaluated by the other half of the warp: groups q (even) and q + 1 sit in one warp
            const int partner = (q % 2 == 0) ? m + 1 : m - 1;
            int L = len_of(m);
            if (partner >= 0 && partner < M && partner / kMelGroups == m / kMelGroups && len_of(partner) > L) L = len_of(partner);
            L = (L + kMelChunk - 1) / kMelChunk * kMelChunk;
            const int first = lo[m] < 0 ? 0 : lo[m];
            // padding taps must read this frame's own (finite) power values: a run that would cross the last bin moves left
            const int start = (first + L <= kBins) ? first : kBins - L;
            AAT_REQUIRE(start >= 0, AAT_ERR_UNSUPPORTED, "aat_create: mel filter %d is wider than the spectrum", m);
            const size_t at = weights.size();
            weights.resize(at + L, 0.0);
            if (lo[m] >= 0)
                for (int k = lo[m]; k <= hi[m]; ++k) weights[at + (k - start)] = filters[(size_t)k * M + m];
            for (int c = 0; c < L / kMelChunk; ++c) {
                const int bin = start + c * kMelChunk;
                if (prev_desc < 0)
                    desc[3 * q + 2] = (uint16_t)bin;
                else
                    desc[prev_desc] |= (uint16_t)bin;
                prev_desc = (int)desc.size();
                desc.push_back(c + 1 == L / kMelChunk ? kMelDescLast : 0);
            }
        }
        desc[3 * q + 0] = (uint16_t)first_chunk;
        desc[3 * q + 1] = (uint16_t)(weights.size() / kMelChunk - first_chunk);
    }
    // the kernel fetches one chunk (and its descriptor) ahead of the one it is adding up
    weights.resize(weights.size() + kMelChunk, 0.0);
    desc.resize(desc.size() + 2, 0);
    AAT_REQUIRE(weights.size() <= (size_t)kMelMaxWeights, AAT_ERR_UNSUPPORTED,
                "aat_create: the mel filter bank is too dense for the log-mel kernel (more than %d banded weights)",
                kMelMaxWeights);
    MelSchedule &ms = ctx->mel;
    ms.n_mels = M;
    ms.nnz = nnz;
    ms.n_weights = (int)weights.size();
    ms.n_desc = (int)desc.size();
    int rc;
    if ((rc = upload(&ms.chunk_desc, desc.data(), desc.size()))) return rc;
    if ((rc = upload(&ms.weight, weights.data(), weights.size()))) return rc;
    return AAT_OK;
}

size_t dtype_size(int dt)
{
    switch (dt) {
    case AAT_F32: return 4;
    case AAT_F64: return 8;
    case AAT_F16:
    case AAT_BF16: return 2;
    default: return 0;
    }
}

} // namespace
} // namespace aat

using namespace aat;

extern "C" {

int aat_version(void) { return AAT_B200_VERSION; }
const char *aat_last_error(void) { return g_error; }
int64_t aat_kernel_launch_count(void) { return g_launch_count.load(); }
void aat_kernel_launch_count_add(int64_t n) { g_launch_count.fetch_add(n); }

int64_t aat_segment_capacity(const aat_config *cfg, int64_t n_samples)
{
    if (!cfg || n_samples < 0 || cfg->max_segment_frames <= 0 || cfg->hop_length <= 0) return -1;
    return seg_capacity(*cfg, n_samples);
}

int64_t aat_num_mel_frames(const aat_config *cfg, int64_t n_samples)
{
    if (!cfg || n_samples < 0 || cfg->hop_length <= 0) return -1;
    return 1 + n_samples / cfg->hop_length;
}

static int init_context(aat_ctx *ctx, const double *window_host, const double *mel_filters_host)
{
    const int device = ctx->device;
    cudaDeviceProp prop{};
    AAT_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;

    // window / 2: the two-frames-per-transform split carries a factor 1/2 (exact power-of-two scaling)
    std::vector<double> win(kNfft);
    for (int i = 0; i < kNfft; ++i) win[i] = 0.5 * window_host[i];
    int rc = upload(&ctx->window_half, win.data(), win.size());
    if (rc) return rc;

    // W_400^(k1 * n2) for k1 = 1, 2, 3, 4, 5, 10, 15 (the kernel derives the other rows as products), long double,
    // rounded once
    const int tw_rows[7] = {1, 2, 3, 4, 5, 10, 15};
    std::vector<double2> tw(7 * 20);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int r = 0; r < 7; ++r)
        for (int n2 = 0; n2 < 20; ++n2) {
            const int e = (tw_rows[r] * n2) % kNfft;
            const long double a = two_pi * (long double)e / (long double)kNfft;
            tw[r * 20 + n2] = make_double2((double)cosl(a), (double)-sinl(a));
        }
    rc = upload(&ctx->twiddle, tw.data(), tw.size());
    if (rc) return rc;

    if ((rc = build_mel_schedule(ctx, mel_filters_host))) return rc;

    if ((rc = logmel_tables_init(ctx))) return rc;
    if ((rc = pool_scratch_init(ctx->num_sms, &ctx->pool))) return rc;
    AAT_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->pool_done, cudaEventDisableTiming));
    AAT_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking));
    return AAT_OK;
}

int aat_create(int device, const aat_config *cfg, const double *window_host, const double *mel_filters_host,
               aat_ctx **out)
{
    AAT_REQUIRE(cfg && window_host && mel_filters_host && out, AAT_ERR_INVALID, "aat_create: NULL argument");
    AAT_REQUIRE(cfg->n_fft == kNfft, AAT_ERR_UNSUPPORTED,
                "aat_create: n_fft=%d is not implemented (the FFT kernel is specialised for n_fft=400)", cfg->n_fft);
    AAT_REQUIRE(cfg->hop_length >= 1 && cfg->hop_length <= kNfft, AAT_ERR_UNSUPPORTED,
                "aat_create: hop_length=%d outside 1..%d", cfg->hop_length, kNfft);
    AAT_REQUIRE(cfg->num_mel_filters >= 1 && cfg->num_mel_filters <= kMaxMels, AAT_ERR_UNSUPPORTED,
                "aat_create: num_mel_filters=%d outside 1..%d", cfg->num_mel_filters, kMaxMels);
    AAT_REQUIRE(cfg->running_mean_points >= 1 && cfg->running_mean_points <= kMaxRunningMean, AAT_ERR_UNSUPPORTED,
                "aat_create: running_mean_points=%d outside 1..%d", cfg->running_mean_points, kMaxRunningMean);
    AAT_REQUIRE(cfg->max_segment_frames > 0, AAT_ERR_INVALID,
                "aat_create: max_segment_frames must be positive (the reference divides by it)");
    AAT_REQUIRE(cfg->min_segment_frames >= 0, AAT_ERR_INVALID, "aat_create: min_segment_frames must be >= 0");
    int n_dev = 0;
    AAT_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
    AAT_REQUIRE(device >= 0 && device < n_dev, AAT_ERR_CUDA, "aat_create: no CUDA device %d (%d visible)", device, n_dev);
    DeviceGuard guard(device);
    AAT_REQUIRE(guard.ok, AAT_ERR_CUDA, "aat_create: cudaSetDevice(%d) failed", device);

    aat_ctx *ctx = new (std::nothrow) aat_ctx();
    AAT_REQUIRE(ctx, AAT_ERR_INVALID, "aat_create: out of host memory");
    ctx->device = device;
    ctx->cfg = *cfg;
    const int rc = init_context(ctx, window_host, mel_filters_host);
    if (rc != AAT_OK) { // aat_destroy copes with a partially built context and keeps the error message
        aat_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return AAT_OK;
}

int aat_destroy(aat_ctx *ctx)
{
    if (!ctx) return AAT_OK;
    DeviceGuard guard(ctx->device);
    cudaFree(ctx->window_half);
    cudaFree(ctx->twiddle);
    cudaFree(ctx->log_table);
    cudaFree(ctx->mel.chunk_desc);
    cudaFree(ctx->mel.weight);
    pool_scratch_free(&ctx->pool);
    if (ctx->pool_done) cudaEventDestroy(ctx->pool_done);
    if (ctx->dev_scratch) cudaFree(ctx->dev_scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->host_stream) cudaStreamDestroy(ctx->host_stream);
    for (auto &kv : ctx->host_plans) aat_plan_destroy(kv.second);
    ctx->host_plans.clear();
    for (cudaEvent_t e : ctx->prof.start) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->prof.stop) cudaEventDestroy(e);
    delete ctx;
    return AAT_OK;
}

int aat_profile_enable(aat_ctx *ctx, uint32_t kernel_mask)
{
    AAT_REQUIRE(ctx, AAT_ERR_INVALID, "aat_profile_enable: NULL context");
    DeviceGuard guard(ctx->device);
    Profiler &p = ctx->prof;
    constexpr size_t kSlots = 16384;
    if (kernel_mask != 0 && p.start.empty()) {
        p.start.resize(kSlots);
        p.stop.resize(kSlots);
        p.kernel.assign(kSlots, 0);
        for (size_t i = 0; i < kSlots; ++i) {
            AAT_CUDA_CHECK(cudaEventCreate(&p.start[i]));
            AAT_CUDA_CHECK(cudaEventCreate(&p.stop[i]));
        }
    }
    p.mask = kernel_mask;
    p.used = 0;
    for (int64_t &n : p.seen) n = 0;
    return AAT_OK;
}

int aat_profile_sample_every(aat_ctx *ctx, int32_t every)
{
    AAT_REQUIRE(ctx && every >= 1, AAT_ERR_INVALID, "aat_profile_sample_every: NULL context or every < 1");
    ctx->prof.every = every;
    return AAT_OK;
}

int aat_profile_summary(aat_ctx *ctx, int64_t *launches, double *total_ms)
{
    AAT_REQUIRE(ctx && launches && total_ms, AAT_ERR_INVALID, "aat_profile_summary: NULL argument");
    DeviceGuard guard(ctx->device);
    Profiler &p = ctx->prof;
    for (int k = 0; k < AAT_K_COUNT; ++k) launches[k] = 0, total_ms[k] = 0.0;
    for (size_t i = 0; i < p.used; ++i) {
        AAT_CUDA_CHECK(cudaEventSynchronize(p.stop[i]));
        float ms = 0.f;
        AAT_CUDA_CHECK(cudaEventElapsedTime(&ms, p.start[i], p.stop[i]));
        launches[p.kernel[i]] += 1;
        total_ms[p.kernel[i]] += ms;
    }
    return AAT_OK;
}

int aat_get_config(const aat_ctx *ctx, aat_config *out)
{
    AAT_REQUIRE(ctx && out, AAT_ERR_INVALID, "aat_get_config: NULL argument");
    *out = ctx->cfg;
    return AAT_OK;
}

// with_pool_scratch: the cross-CTA scratch of the pool kernel (14.5 MB); the single-utterance plans cached for the
// aat_host_* entry points never pool and go without.
static int plan_create(aat_ctx *ctx, int32_t n_utts, const int64_t *n_samples_host, bool with_pool_scratch, aat_plan **out)
{
    AAT_REQUIRE(ctx && out && (n_samples_host || n_utts == 0), AAT_ERR_INVALID, "aat_plan_create: NULL argument");
    AAT_REQUIRE(n_utts >= 0, AAT_ERR_INVALID, "aat_plan_create: negative n_utts");
    DeviceGuard guard(ctx->device);
    aat_plan *plan = new (std::nothrow) aat_plan();
    AAT_REQUIRE(plan, AAT_ERR_INVALID, "aat_plan_create: out of host memory");
    plan->ctx = ctx;
    plan->n_utts = n_utts;
    plan->h_n_samples.assign(n_samples_host, n_samples_host + n_utts);
    plan->h_wave_off.assign(n_utts + 1, 0);
    plan->h_frame_off.assign(n_utts + 1, 0);
    plan->h_seg_slot_off.assign(n_utts + 1, 0);
    std::vector<int32_t> chunk_first(n_utts + 1, 0), chunk_utt;
    std::vector<int64_t> burst_off(n_utts + 1, 0);
    std::vector<MelTile> tiles_desc;
    const int hop = ctx->cfg.hop_length, n_mels = ctx->cfg.num_mel_filters;
    const int64_t stage_pad = (((int64_t)(kMelFramesPerTile - 1) * hop + kNfft) + 3) & ~int64_t(3);
    for (int b = 0; b < n_utts; ++b) {
        const int64_t n = n_samples_host[b];
        if (n < 1) {
            delete plan;
            AAT_REQUIRE(false, AAT_ERR_INVALID,
                        "aat_plan_create: utterance %d has %lld samples (np.pad(mode='reflect') needs >= 1)", b,
                        (long long)n);
        }
        const int64_t T = 1 + n / ctx->cfg.hop_length;
        plan->h_wave_off[b + 1] = plan->h_wave_off[b] + n;
        plan->h_frame_off[b + 1] = plan->h_frame_off[b] + T;
        plan->h_seg_slot_off[b + 1] = plan->h_seg_slot_off[b] + seg_capacity(ctx->cfg, n);
        if (T > plan->max_frames) plan->max_frames = T;
        const int64_t tiles = (T + kMelFramesPerTile - 1) / kMelFramesPerTile;
        if ((int64_t)tiles_desc.size() + tiles > (int64_t)INT32_MAX || T > (int64_t)INT32_MAX) {
            delete plan;
            AAT_REQUIRE(false, AAT_ERR_UNSUPPORTED, "aat_plan_create: batch too large (mel tiles exceed 2^31)");
        }
        for (int64_t t = 0; t < tiles; ++t) {
            const int64_t f0 = t * kMelFramesPerTile, g0 = f0 * hop - kNfft / 2;
            MelTile d{};
            d.src = plan->h_wave_off[b] + g0;
            d.n = n;
            d.wave_off = plan->h_wave_off[b];
            d.mel_off = (int64_t)n_mels * plan->h_frame_off[b] + f0;
            d.amp_off = plan->h_frame_off[b] + f0;
            d.T = (int32_t)T;
            d.valid = (int32_t)(T - f0 < kMelFramesPerTile ? T - f0 : kMelFramesPerTile);
            d.interior = (g0 >= 0 && g0 + stage_pad <= n) ? 1 : 0;
            d.utt = b;
            tiles_desc.push_back(d);
        }
        burst_off[b + 1] = burst_off[b] + synth_burst_capacity(ctx->cfg.sampling_rate, n);
        const int64_t chunks = (n + kNormChunk - 1) / kNormChunk;
        chunk_first[b + 1] = chunk_first[b] + (int32_t)chunks;
        chunk_utt.insert(chunk_utt.end(), (size_t)chunks, b);
    }
    plan->total_samples = plan->h_wave_off[n_utts];
    plan->total_frames = plan->h_frame_off[n_utts];
    plan->total_seg_slots = plan->h_seg_slot_off[n_utts];
    plan->mel_tiles = (int32_t)tiles_desc.size();
    int rc;
    if ((rc = upload(&plan->d_n_samples, plan->h_n_samples.data(), (size_t)n_utts)) ||
        (rc = upload(&plan->d_wave_off, plan->h_wave_off.data(), (size_t)n_utts + 1)) ||
        (rc = upload(&plan->d_frame_off, plan->h_frame_off.data(), (size_t)n_utts + 1)) ||
        (rc = upload(&plan->d_seg_slot_off, plan->h_seg_slot_off.data(), (size_t)n_utts + 1)) ||
        (rc = upload(&plan->d_mel_tile, tiles_desc.data(), tiles_desc.size())) ||
        (rc = upload(&plan->d_chunk_utt, chunk_utt.data(), chunk_utt.size())) ||
        (rc = upload(&plan->d_chunk_first, chunk_first.data(), chunk_first.size())) ||
        (rc = upload(&plan->d_burst_off, burst_off.data(), burst_off.size()))) {
        aat_plan_destroy(plan);
        return rc;
    }
    plan->norm_chunks = (int32_t)chunk_utt.size();
    plan->total_bursts = burst_off[n_utts];
    if (cudaMalloc(&plan->d_norm_partial, sizeof(double) * 3 * (size_t)(plan->norm_chunks ? plan->norm_chunks : 1)) != cudaSuccess ||
        cudaMalloc(&plan->d_norm_stats, sizeof(double) * 2 * (size_t)(n_utts ? n_utts : 1)) != cudaSuccess) {
        aat_plan_destroy(plan);
        AAT_REQUIRE(false, AAT_ERR_CUDA, "aat_plan_create: out of device memory");
    }
    if (cudaMalloc(&plan->d_mel_sched, sizeof(int32_t) * 4) != cudaSuccess ||
        cudaMemset(plan->d_mel_sched, 0, sizeof(int32_t) * 4) != cudaSuccess) {
        aat_plan_destroy(plan);
        AAT_REQUIRE(false, AAT_ERR_CUDA, "aat_plan_create: out of device memory");
    }
    if (cudaMalloc(&plan->d_seg_local, sizeof(int64_t) * (size_t)(plan->total_seg_slots ? plan->total_seg_slots : 1)) != cudaSuccess ||
        cudaMalloc(&plan->d_utt_frames, sizeof(int64_t) * (size_t)(n_utts ? n_utts : 1)) != cudaSuccess ||
        cudaMemset(plan->d_utt_frames, 0, sizeof(int64_t) * (size_t)(n_utts ? n_utts : 1)) != cudaSuccess) {
        aat_plan_destroy(plan);
        AAT_REQUIRE(false, AAT_ERR_CUDA, "aat_plan_create: out of device memory");
    }
    if (with_pool_scratch) {
        const int prc = pool_scratch_init(ctx->num_sms, &plan->pool);
        if (prc != AAT_OK) {
            aat_plan_destroy(plan);
            return prc;
        }
    }
    *out = plan;
    return AAT_OK;
}

int aat_plan_create(aat_ctx *ctx, int32_t n_utts, const int64_t *n_samples_host, aat_plan **out)
{
    return plan_create(ctx, n_utts, n_samples_host, true, out);
}

int aat_plan_destroy(aat_plan *plan)
{
    if (!plan) return AAT_OK;
    DeviceGuard guard(plan->ctx->device);
    cudaFree(plan->d_n_samples);
    cudaFree(plan->d_wave_off);
    cudaFree(plan->d_frame_off);
    cudaFree(plan->d_seg_slot_off);
    cudaFree(plan->d_mel_tile);
    cudaFree(plan->d_mel_sched);
    cudaFree(plan->d_seg_local);
    cudaFree(plan->d_utt_frames);
    cudaFree(plan->d_chunk_utt);
    cudaFree(plan->d_chunk_first);
    cudaFree(plan->d_norm_partial);
    cudaFree(plan->d_norm_stats);
    cudaFree(plan->d_burst_off);
    pool_scratch_free(&plan->pool);
    delete plan;
    return AAT_OK;
}

int64_t aat_plan_total_samples(const aat_plan *plan) { return plan ? plan->total_samples : -1; }
int64_t aat_plan_total_frames(const aat_plan *plan) { return plan ? plan->total_frames : -1; }
int64_t aat_plan_total_seg_slots(const aat_plan *plan) { return plan ? plan->total_seg_slots : -1; }

int aat_plan_offsets(const aat_plan *plan, int64_t *wave_off_host, int64_t *frame_off_host, int64_t *seg_slot_off_host)
{
    AAT_REQUIRE(plan, AAT_ERR_INVALID, "aat_plan_offsets: NULL plan");
    const size_t bytes = sizeof(int64_t) * ((size_t)plan->n_utts + 1);
    if (wave_off_host) memcpy(wave_off_host, plan->h_wave_off.data(), bytes);
    if (frame_off_host) memcpy(frame_off_host, plan->h_frame_off.data(), bytes);
    if (seg_slot_off_host) memcpy(seg_slot_off_host, plan->h_seg_slot_off.data(), bytes);
    return AAT_OK;
}

// ------------------------------------------------------------------------------------------------ device API
int aat_logmel(aat_ctx *ctx, const aat_plan *plan, const void *wave_dev, int wave_dtype, const double *znorm_stats_dev,
               float *mel_dev, float *amp_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && wave_dev && mel_dev, AAT_ERR_INVALID, "aat_logmel: NULL argument");
    AAT_REQUIRE(plan->ctx == ctx, AAT_ERR_INVALID, "aat_logmel: plan belongs to another context");
    AAT_DEVICE_GUARD(ctx);
    return launch_logmel(ctx, plan, wave_dev, wave_dtype, znorm_stats_dev, mel_dev, amp_dev,
                         static_cast<cudaStream_t>(stream));
}

int aat_amplitude(aat_ctx *ctx, const aat_plan *plan, const float *mel_dev, float *amp_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && mel_dev && amp_dev, AAT_ERR_INVALID, "aat_amplitude: NULL argument");
    AAT_REQUIRE(plan->ctx == ctx, AAT_ERR_INVALID, "aat_amplitude: plan belongs to another context");
    AAT_REQUIRE(plan->n_utts <= 65535, AAT_ERR_UNSUPPORTED, "aat_amplitude: at most 65535 utterances per plan");
    AAT_DEVICE_GUARD(ctx);
    return launch_amplitude(ctx, plan, mel_dev, amp_dev, static_cast<cudaStream_t>(stream));
}

int aat_boundaries(aat_ctx *ctx, const aat_plan *plan, const float *mel_dev, const float *amp_dev,
                   int64_t *seg_start_dev, int64_t *seg_len_dev, int32_t *seg_count_dev, int64_t *minima_dev,
                   int32_t *minima_count_dev, int32_t *status_dev, int64_t *seg_off_dev, int64_t *n_seg_dev,
                   int64_t *utt_seg_off_dev, void *stream)
{
    AAT_REQUIRE(seg_off_dev == nullptr || n_seg_dev != nullptr, AAT_ERR_INVALID,
                "aat_boundaries: seg_off_dev needs n_seg_dev");
    AAT_REQUIRE(ctx && plan && seg_start_dev && seg_len_dev && seg_count_dev && status_dev, AAT_ERR_INVALID,
                "aat_boundaries: NULL argument");
    AAT_REQUIRE(mel_dev || amp_dev, AAT_ERR_INVALID, "aat_boundaries: need mel_dev or amp_dev");
    AAT_REQUIRE(plan->ctx == ctx, AAT_ERR_INVALID, "aat_boundaries: plan belongs to another context");
    AAT_DEVICE_GUARD(ctx);
    return launch_boundaries(ctx, plan, mel_dev, amp_dev, seg_start_dev, seg_len_dev, seg_count_dev, minima_dev,
                             minima_count_dev, status_dev, seg_off_dev, n_seg_dev, utt_seg_off_dev,
                             static_cast<cudaStream_t>(stream));
}

int aat_process_boarders(aat_ctx *ctx, int64_t n_samples, const int64_t *boarders_dev, int64_t n_boarders,
                         int64_t *seg_start_dev, int64_t *seg_len_dev, int64_t capacity, int32_t *seg_count_dev,
                         int32_t *status_dev, void *stream)
{
    AAT_REQUIRE(ctx && (boarders_dev || n_boarders == 0) && seg_start_dev && seg_len_dev && seg_count_dev && status_dev,
                AAT_ERR_INVALID, "aat_process_boarders: NULL argument");
    AAT_REQUIRE(n_samples >= 0 && n_boarders >= 0 && capacity >= 0, AAT_ERR_INVALID, "aat_process_boarders: negative size");
    AAT_DEVICE_GUARD(ctx);
    return launch_process_boarders(ctx, n_samples, boarders_dev, n_boarders, seg_start_dev, seg_len_dev, capacity,
                                   seg_count_dev, status_dev, static_cast<cudaStream_t>(stream));
}

int aat_segment_frame_csr(aat_ctx *ctx, const aat_plan *plan, const int64_t *seg_len_dev, const int32_t *seg_count_dev,
                          int64_t *seg_off_dev, int64_t *n_seg_dev, int64_t *utt_seg_off_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && seg_len_dev && seg_count_dev && seg_off_dev && n_seg_dev, AAT_ERR_INVALID,
                "aat_segment_frame_csr: NULL argument");
    AAT_DEVICE_GUARD(ctx);
    return launch_segment_frame_csr(ctx, plan, seg_len_dev, seg_count_dev, seg_off_dev, n_seg_dev, utt_seg_off_dev,
                                    static_cast<cudaStream_t>(stream));
}

int aat_utterance_frame_csr(aat_ctx *ctx, const aat_plan *plan, const int64_t *seg_start_dev,
                            const int32_t *seg_count_dev, const int64_t *utt_seg_off_dev, int64_t *seg_off_dev,
                            int64_t *n_seg_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && seg_start_dev && seg_count_dev && utt_seg_off_dev && seg_off_dev && n_seg_dev,
                AAT_ERR_INVALID, "aat_utterance_frame_csr: NULL argument");
    AAT_REQUIRE(plan->ctx == ctx, AAT_ERR_INVALID, "aat_utterance_frame_csr: plan belongs to another context");
    AAT_DEVICE_GUARD(ctx);
    return launch_utterance_frame_csr(plan, seg_start_dev, seg_count_dev, utt_seg_off_dev, seg_off_dev, n_seg_dev,
                                      static_cast<cudaStream_t>(stream));
}

int aat_segment_mean_pool(aat_ctx *ctx, const aat_plan *plan, const void *emb_dev, int emb_dtype, int64_t n_rows,
                          int32_t dim, const int64_t *seg_off_dev, int64_t n_seg, const int64_t *n_seg_dev,
                          float *out_dev, double *colsum_dev, int flags, void *stream)
{
    AAT_REQUIRE(ctx, AAT_ERR_INVALID, "aat_segment_mean_pool: NULL context");
    AAT_REQUIRE(plan == nullptr || plan->ctx == ctx, AAT_ERR_INVALID, "aat_segment_mean_pool: plan belongs to another context");
    AAT_REQUIRE((flags & ~(AAT_POOL_ACCUMULATE | AAT_POOL_EMB_READY | AAT_POOL_ROWS_FROM_DEVICE | AAT_POOL_SHARE_SMS)) == 0, AAT_ERR_INVALID,
                "aat_segment_mean_pool: unknown flag bits 0x%x", flags);
    AAT_DEVICE_GUARD(ctx);
    return launch_mean_pool(ctx, plan, emb_dev, emb_dtype, n_rows, dim, seg_off_dev, n_seg, n_seg_dev, out_dev,
                            colsum_dev, flags, static_cast<cudaStream_t>(stream));
}

int aat_tokenize_and_pool(aat_ctx *ctx, const aat_plan *plan, const aat_step_buffers *bufs, const void *wave_dev,
                          int wave_dtype, int znorm, const void *emb_dev, int emb_dtype, int64_t n_rows, int32_t dim,
                          float *out_dev, int64_t out_capacity, double *colsum_dev, int pool_flags, void *stream)
{
    AAT_REQUIRE(ctx && plan && bufs && wave_dev && out_dev, AAT_ERR_INVALID, "aat_tokenize_and_pool: NULL argument");
    AAT_REQUIRE(plan->ctx == ctx, AAT_ERR_INVALID, "aat_tokenize_and_pool: plan belongs to another context");
    AAT_REQUIRE(bufs->mel && bufs->amp && bufs->seg_start && bufs->seg_len && bufs->seg_count && bufs->status &&
                    bufs->seg_off && bufs->n_seg,
                AAT_ERR_INVALID, "aat_tokenize_and_pool: a required step buffer is NULL");
    AAT_REQUIRE(!znorm || bufs->znorm_stats, AAT_ERR_INVALID, "aat_tokenize_and_pool: znorm needs bufs->znorm_stats");
    AAT_REQUIRE((pool_flags & ~(AAT_POOL_ACCUMULATE | AAT_POOL_EMB_READY | AAT_POOL_ROWS_FROM_DEVICE | AAT_POOL_SHARE_SMS)) == 0, AAT_ERR_INVALID,
                "aat_tokenize_and_pool: unknown flag bits 0x%x", pool_flags);
    AAT_DEVICE_GUARD(ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (znorm && (rc = launch_normalize(ctx, plan, wave_dev, wave_dtype, AAT_NORM_ZSCORE, nullptr, AAT_F64, bufs->znorm_stats, st)))
        return rc;
    if ((rc = launch_logmel(ctx, plan, wave_dev, wave_dtype, znorm ? bufs->znorm_stats : nullptr, bufs->mel, bufs->amp, st)))
        return rc;
    if ((rc = launch_boundaries(ctx, plan, bufs->mel, bufs->amp, bufs->seg_start, bufs->seg_len, bufs->seg_count, bufs->minima,
                                bufs->minima_count, bufs->status, bufs->seg_off, bufs->n_seg, bufs->utt_seg_off, st)))
        return rc;
    return launch_mean_pool(ctx, plan, emb_dev, emb_dtype, n_rows, dim, bufs->seg_off, out_capacity, bufs->n_seg, out_dev,
                            colsum_dev, pool_flags, st);
}

int aat_colsum_accumulate(aat_ctx *ctx, double *acc_dev, const double *colsum_dev, int32_t dim, void *stream)
{
    AAT_REQUIRE(ctx && acc_dev && colsum_dev && dim > 0, AAT_ERR_INVALID, "aat_colsum_accumulate: bad argument");
    AAT_DEVICE_GUARD(ctx);
    return launch_colsum_accumulate(acc_dev, colsum_dev, dim, static_cast<cudaStream_t>(stream));
}

int aat_colsum_finalize(aat_ctx *ctx, const double *acc_dev, int32_t dim, float *mean_dev, void *stream)
{
    AAT_REQUIRE(ctx && acc_dev && mean_dev && dim > 0, AAT_ERR_INVALID, "aat_colsum_finalize: bad argument");
    AAT_DEVICE_GUARD(ctx);
    return launch_colsum_finalize(acc_dev, dim, mean_dev, static_cast<cudaStream_t>(stream));
}

int aat_normalize(aat_ctx *ctx, const aat_plan *plan, const void *wave_dev, int in_dtype, int mode, void *out_dev,
                  int out_dtype, double *stats_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && wave_dev && (out_dev || stats_dev), AAT_ERR_INVALID, "aat_normalize: NULL argument");
    AAT_REQUIRE(plan->ctx == ctx, AAT_ERR_INVALID, "aat_normalize: plan belongs to another context");
    AAT_DEVICE_GUARD(ctx);
    return launch_normalize(ctx, plan, wave_dev, in_dtype, mode, out_dev, out_dtype, stats_dev,
                            static_cast<cudaStream_t>(stream));
}

int aat_pad_segment_boarders(aat_ctx *ctx, const aat_plan *plan, const int64_t *seg_len_dev,
                             const int32_t *seg_count_dev, int64_t s_max, int64_t *boarders_dev, int64_t *mask_dev,
                             int32_t *status_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && seg_len_dev && seg_count_dev && boarders_dev && mask_dev && status_dev, AAT_ERR_INVALID,
                "aat_pad_segment_boarders: NULL argument");
    AAT_REQUIRE(s_max >= 0, AAT_ERR_INVALID, "aat_pad_segment_boarders: negative s_max");
    AAT_DEVICE_GUARD(ctx);
    return launch_pad_boarders(plan, seg_len_dev, seg_count_dev, s_max, boarders_dev, mask_dev, status_dev,
                               static_cast<cudaStream_t>(stream));
}

int aat_scatter_segments(aat_ctx *ctx, const float *wave_padded_dev, int64_t n_max, int32_t n_utts,
                         const int64_t *boarders_dev, int64_t s_max, int64_t max_frames, float *out_dev,
                         float *mask_dev, int32_t *status_dev, void *stream)
{
    AAT_REQUIRE(ctx && wave_padded_dev && boarders_dev && out_dev && status_dev, AAT_ERR_INVALID,
                "aat_scatter_segments: NULL argument");
    AAT_REQUIRE(n_max >= 0 && n_utts >= 0 && s_max >= 0 && max_frames >= 0, AAT_ERR_INVALID,
                "aat_scatter_segments: negative size");
    AAT_DEVICE_GUARD(ctx);
    return launch_scatter_segments(wave_padded_dev, n_max, n_utts, boarders_dev, s_max, max_frames, out_dev, mask_dev,
                                   status_dev, static_cast<cudaStream_t>(stream));
}

int aat_scatter_mel_segments(aat_ctx *ctx, const aat_plan *plan, const float *mel_dev, const int64_t *boarders_dev,
                             int64_t s_max, int64_t max_items, float *out_dev, int32_t *status_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && mel_dev && boarders_dev && out_dev && status_dev, AAT_ERR_INVALID,
                "aat_scatter_mel_segments: NULL argument");
    AAT_REQUIRE(s_max >= 0 && max_items >= 0, AAT_ERR_INVALID, "aat_scatter_mel_segments: negative size");
    AAT_DEVICE_GUARD(ctx);
    return launch_scatter_mel_segments(ctx, plan, plan->n_utts, mel_dev, nullptr, nullptr, nullptr, boarders_dev, s_max,
                                       max_items, out_dev, status_dev, static_cast<cudaStream_t>(stream));
}

int aat_scatter_mel_tiles(aat_ctx *ctx, int32_t n_utts, const float *mel_dev, const int64_t *mel_elem_off_dev,
                          const int64_t *mel_frames_dev, const int64_t *mel_row_stride_dev, const int64_t *boarders_dev,
                          int64_t s_max, int64_t max_items, float *out_dev, int32_t *status_dev, void *stream)
{
    AAT_REQUIRE(ctx && mel_dev && mel_elem_off_dev && mel_frames_dev && mel_row_stride_dev && boarders_dev && out_dev &&
                    status_dev,
                AAT_ERR_INVALID, "aat_scatter_mel_tiles: NULL argument");
    AAT_REQUIRE(n_utts >= 0 && s_max >= 0 && max_items >= 0, AAT_ERR_INVALID, "aat_scatter_mel_tiles: negative size");
    AAT_DEVICE_GUARD(ctx);
    return launch_scatter_mel_segments(ctx, nullptr, n_utts, mel_dev, mel_elem_off_dev, mel_frames_dev, mel_row_stride_dev,
                                       boarders_dev, s_max, max_items, out_dev, status_dev,
                                       static_cast<cudaStream_t>(stream));
}

int aat_normalize_padded(aat_ctx *ctx, const aat_plan *plan, const void *wave_dev, int in_dtype, int mode, float *out_dev,
                         int64_t n_max, int32_t *mask_dev, double *stats_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && wave_dev && out_dev, AAT_ERR_INVALID, "aat_normalize_padded: NULL argument");
    AAT_REQUIRE(plan->ctx == ctx, AAT_ERR_INVALID, "aat_normalize_padded: plan belongs to another context");
    AAT_REQUIRE(n_max >= 0, AAT_ERR_INVALID, "aat_normalize_padded: negative n_max");
    for (int64_t n : plan->h_n_samples)
        AAT_REQUIRE(n <= n_max, AAT_ERR_INVALID, "aat_normalize_padded: an utterance of %lld samples does not fit n_max = %lld",
                    (long long)n, (long long)n_max);
    AAT_DEVICE_GUARD(ctx);
    return launch_normalize_padded(ctx, plan, wave_dev, in_dtype, mode, out_dev, n_max, mask_dev, stats_dev,
                                   static_cast<cudaStream_t>(stream));
}

int aat_masked_mean_pool(aat_ctx *ctx, const void *emb_dev, int emb_dtype, int64_t n_rows, int64_t seq_len, int32_t dim,
                         const int64_t *mask_dev, float *out_dev, int64_t *row_mask_dev, void *stream)
{
    AAT_REQUIRE(ctx && (emb_dev || n_rows == 0) && mask_dev && out_dev, AAT_ERR_INVALID, "aat_masked_mean_pool: NULL argument");
    AAT_REQUIRE(n_rows >= 0 && seq_len >= 0 && dim > 0, AAT_ERR_INVALID, "aat_masked_mean_pool: negative size");
    AAT_DEVICE_GUARD(ctx);
    return launch_masked_mean_pool(emb_dev, emb_dtype, n_rows, seq_len, dim, mask_dev, out_dev, row_mask_dev,
                                   static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------------ synthetic inputs
int64_t aat_synth_workspace_bytes(const aat_plan *plan) { return plan ? (int64_t)synth_workspace_bytes(plan) : -1; }

int aat_synth_waveforms(aat_ctx *ctx, const aat_plan *plan, uint64_t seed_base, int64_t utt_index_base, float *wave_dev,
                        void *workspace_dev, void *stream)
{
    AAT_REQUIRE(ctx && plan && wave_dev && workspace_dev, AAT_ERR_INVALID, "aat_synth_waveforms: NULL argument");
    AAT_REQUIRE(plan->ctx == ctx, AAT_ERR_INVALID, "aat_synth_waveforms: plan belongs to another context");
    AAT_DEVICE_GUARD(ctx);
    return launch_synth_waveforms(ctx, plan, seed_base, utt_index_base, wave_dev, workspace_dev,
                                  static_cast<cudaStream_t>(stream));
}

int aat_synth_normal(aat_ctx *ctx, float *out_dev, int64_t n, uint64_t seed, void *stream)
{
    AAT_REQUIRE(ctx && (out_dev || n == 0) && n >= 0, AAT_ERR_INVALID, "aat_synth_normal: bad argument");
    AAT_DEVICE_GUARD(ctx);
    return launch_synth_normal(ctx, out_dev, n, seed, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------------ host API
// One utterance, host buffers in / host buffers out: H2D -> kernels -> D2H on the context's own stream.

// Single-utterance plan for `n_samples`, from the context's cache (most recently used first, 8 entries).
static int host_plan_for(aat_ctx *ctx, int64_t n_samples, aat_plan **out)
{
    auto &cache = ctx->host_plans;
    for (size_t i = 0; i < cache.size(); ++i)
        if (cache[i].first == n_samples) {
            auto hit = cache[i];
            cache.erase(cache.begin() + (long)i);
            cache.insert(cache.begin(), hit);
            *out = hit.second;
            return AAT_OK;
        }
    aat_plan *plan = nullptr;
    int rc = plan_create(ctx, 1, &n_samples, false, &plan);
    if (rc) return rc;
    cache.insert(cache.begin(), std::make_pair(n_samples, plan));
    if (cache.size() > 8) {
        aat_plan_destroy(cache.back().second);
        cache.pop_back();
    }
    *out = plan;
    return AAT_OK;
}

static int host_pipeline(aat_ctx *ctx, const void *wave_host, int wave_dtype, int64_t n_samples,
                         const float *mel_in_host, bool want_boundaries, float *mel_out_host, int64_t *minima_host,
                         int64_t *n_minima_host, int64_t *seg_start_host, int64_t *seg_len_host, int64_t capacity,
                         int64_t *n_segments_host, int32_t *padded_tail_host)
{
    DeviceGuard guard(ctx->device);
    AAT_REQUIRE(guard.ok, AAT_ERR_CUDA, "cudaSetDevice(%d) failed", ctx->device);
    std::lock_guard<std::mutex> lock(ctx->host_mutex);
    aat_plan *plan = nullptr;
    int rc = host_plan_for(ctx, n_samples, &plan);
    if (rc) return rc;
    const int M = ctx->cfg.num_mel_filters;
    const int64_t T = plan->total_frames;
    const int64_t slots = plan->total_seg_slots;
    const size_t wsize = dtype_size(wave_dtype);
    const bool need_wave = mel_in_host == nullptr;

    size_t dev_bytes = align256(need_wave ? wsize * n_samples : 0) + align256(sizeof(float) * M * T) +
                       align256(sizeof(float) * T) + 2 * align256(sizeof(int64_t) * slots) +
                       align256(sizeof(int64_t) * T) + 3 * 256;
    size_t pin_bytes = dev_bytes;
    rc = ensure_scratch(ctx, dev_bytes, pin_bytes);
    if (rc) return rc;
    Arena d(ctx->dev_scratch), h(ctx->pinned);
    unsigned char *d_wave = d.take<unsigned char>(need_wave ? wsize * n_samples : 0);
    float *d_mel = d.take<float>((size_t)M * T);
    float *d_amp = d.take<float>(T);
    int64_t *d_seg_start = d.take<int64_t>(slots);
    int64_t *d_seg_len = d.take<int64_t>(slots);
    int64_t *d_minima = d.take<int64_t>(T);
    int32_t *d_seg_count = d.take<int32_t>(1);
    int32_t *d_min_count = d.take<int32_t>(1);
    int32_t *d_status = d.take<int32_t>(1);
    unsigned char *h_wave = h.take<unsigned char>(need_wave ? wsize * n_samples : 0);
    float *h_mel = h.take<float>((size_t)M * T);
    (void)h.take<float>(T);
    int64_t *h_seg_start = h.take<int64_t>(slots);
    int64_t *h_seg_len = h.take<int64_t>(slots);
    int64_t *h_minima = h.take<int64_t>(T);
    int32_t *h_seg_count = h.take<int32_t>(1);
    int32_t *h_min_count = h.take<int32_t>(1);
    int32_t *h_status = h.take<int32_t>(1);
    cudaStream_t st = ctx->host_stream;

    auto fail = [&](int code) { return code; }; // the plan stays in the context's cache
#define AAT_TRY_CUDA(expr)                                                                            \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);   \
            return fail(AAT_ERR_CUDA);                                                                \
        }                                                                                             \
    } while (0)

    if (need_wave) {
        memcpy(h_wave, wave_host, wsize * n_samples);
        AAT_TRY_CUDA(cudaMemcpyAsync(d_wave, h_wave, wsize * n_samples, cudaMemcpyHostToDevice, st));
        rc = launch_logmel(ctx, plan, d_wave, wave_dtype, nullptr, d_mel, want_boundaries ? d_amp : nullptr, st);
        if (rc) return fail(rc);
        if (mel_out_host) AAT_TRY_CUDA(cudaMemcpyAsync(h_mel, d_mel, sizeof(float) * M * T, cudaMemcpyDeviceToHost, st));
    } else {
        memcpy(h_mel, mel_in_host, sizeof(float) * M * T);
        AAT_TRY_CUDA(cudaMemcpyAsync(d_mel, h_mel, sizeof(float) * M * T, cudaMemcpyHostToDevice, st));
    }
    if (want_boundaries) {
        rc = launch_boundaries(ctx, plan, d_mel, need_wave ? d_amp : nullptr, d_seg_start, d_seg_len, d_seg_count,
                               d_minima, d_min_count, d_status, nullptr, nullptr, nullptr, st);
        if (rc) return fail(rc);
        // segment tables, minima and the three counters sit back to back in both arenas (same layout): ONE copy
        // (each cudaMemcpyAsync costs 3-5 us of host time, and this path is latency-bound)
        const size_t span = (size_t)(reinterpret_cast<unsigned char *>(d_status) - reinterpret_cast<unsigned char *>(d_seg_start)) +
                            sizeof(int32_t);
        AAT_TRY_CUDA(cudaMemcpyAsync(h_seg_start, d_seg_start, span, cudaMemcpyDeviceToHost, st));
    }
    AAT_TRY_CUDA(cudaStreamSynchronize(st));
#undef AAT_TRY_CUDA
    if (mel_out_host) memcpy(mel_out_host, h_mel, sizeof(float) * M * T);
    if (want_boundaries) {
        if (*h_status < 0) {
            set_error(*h_status == AAT_ERR_TAIL ? "tail longer than min_segment_frames (reference raises ValueError)"
                                                : "segment capacity exceeded");
            return fail(*h_status);
        }
        if (padded_tail_host) *padded_tail_host = *h_status & 1;
        const int64_t n_seg = *h_seg_count;
        if (n_seg > capacity) {
            set_error("aat_host_tokenize: %lld segments exceed the caller's capacity %lld", (long long)n_seg,
                      (long long)capacity);
            return fail(AAT_ERR_CAPACITY);
        }
        if (seg_start_host) memcpy(seg_start_host, h_seg_start, sizeof(int64_t) * n_seg);
        if (seg_len_host) memcpy(seg_len_host, h_seg_len, sizeof(int64_t) * n_seg);
        if (n_segments_host) *n_segments_host = n_seg;
        if (minima_host) memcpy(minima_host, h_minima, sizeof(int64_t) * (*h_min_count));
        if (n_minima_host) *n_minima_host = *h_min_count;
    }
    return AAT_OK;
}

int aat_host_logmel(aat_ctx *ctx, const void *wave_host, int wave_dtype, int64_t n_samples, float *mel_host)
{
    AAT_REQUIRE(ctx && wave_host && mel_host, AAT_ERR_INVALID, "aat_host_logmel: NULL argument");
    AAT_REQUIRE(wave_dtype == AAT_F32 || wave_dtype == AAT_F64, AAT_ERR_UNSUPPORTED,
                "aat_host_logmel: waveform dtype must be AAT_F32 or AAT_F64");
    return host_pipeline(ctx, wave_host, wave_dtype, n_samples, nullptr, false, mel_host, nullptr, nullptr, nullptr,
                         nullptr, 0, nullptr, nullptr);
}

int aat_host_find_minimas(aat_ctx *ctx, const float *mel_host, int64_t n_frames, int64_t *minima_host,
                          int64_t *n_minima_host)
{
    AAT_REQUIRE(ctx && mel_host && minima_host && n_minima_host, AAT_ERR_INVALID, "aat_host_find_minimas: NULL argument");
    AAT_REQUIRE(n_frames >= 1, AAT_ERR_INVALID, "aat_host_find_minimas: need at least one frame");
    // any sample count with 1 + n/hop == n_frames reproduces the layout; the segments are discarded
    const int64_t n_samples = (n_frames - 1) * ctx->cfg.hop_length + (n_frames == 1 ? 1 : 0);
    std::vector<int64_t> seg((size_t)seg_capacity(ctx->cfg, n_samples) * 2);
    int64_t n_seg = 0;
    return host_pipeline(ctx, nullptr, AAT_F32, n_samples, mel_host, true, nullptr, minima_host, n_minima_host,
                         seg.data(), seg.data() + seg.size() / 2, (int64_t)seg.size() / 2, &n_seg, nullptr);
}

int aat_host_process_boarders(aat_ctx *ctx, int64_t n_samples, const int64_t *boarders_host, int64_t n_boarders,
                              int64_t *seg_start_host, int64_t *seg_len_host, int64_t capacity, int64_t *n_segments_host,
                              int32_t *padded_tail_host)
{
    AAT_REQUIRE(ctx && (boarders_host || n_boarders == 0) && seg_start_host && seg_len_host && n_segments_host,
                AAT_ERR_INVALID, "aat_host_process_boarders: NULL argument");
    AAT_REQUIRE(n_samples >= 0 && n_boarders >= 0 && capacity >= 0, AAT_ERR_INVALID,
                "aat_host_process_boarders: negative size");
    DeviceGuard guard(ctx->device);
    std::lock_guard<std::mutex> lock(ctx->host_mutex);
    size_t bytes = align256(sizeof(int64_t) * n_boarders) + 2 * align256(sizeof(int64_t) * capacity) + 2 * 256;
    int rc = ensure_scratch(ctx, bytes, bytes);
    if (rc) return rc;
    Arena d(ctx->dev_scratch), h(ctx->pinned);
    int64_t *d_b = d.take<int64_t>(n_boarders), *d_s = d.take<int64_t>(capacity), *d_l = d.take<int64_t>(capacity);
    int32_t *d_cnt = d.take<int32_t>(1), *d_status = d.take<int32_t>(1);
    int64_t *h_b = h.take<int64_t>(n_boarders), *h_s = h.take<int64_t>(capacity), *h_l = h.take<int64_t>(capacity);
    int32_t *h_cnt = h.take<int32_t>(1), *h_status = h.take<int32_t>(1);
    cudaStream_t st = ctx->host_stream;
    if (n_boarders) memcpy(h_b, boarders_host, sizeof(int64_t) * n_boarders);
    AAT_CUDA_CHECK(cudaMemcpyAsync(d_b, h_b, sizeof(int64_t) * n_boarders, cudaMemcpyHostToDevice, st));
    rc = launch_process_boarders(ctx, n_samples, d_b, n_boarders, d_s, d_l, capacity, d_cnt, d_status, st);
    if (rc) return rc;
    AAT_CUDA_CHECK(cudaMemcpyAsync(h_s, d_s, sizeof(int64_t) * capacity, cudaMemcpyDeviceToHost, st));
    AAT_CUDA_CHECK(cudaMemcpyAsync(h_l, d_l, sizeof(int64_t) * capacity, cudaMemcpyDeviceToHost, st));
    AAT_CUDA_CHECK(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    AAT_CUDA_CHECK(cudaMemcpyAsync(h_status, d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    AAT_CUDA_CHECK(cudaStreamSynchronize(st));
    if (*h_status < 0) {
        set_error(*h_status == AAT_ERR_TAIL ? "tail longer than min_segment_frames (reference raises ValueError)"
                                            : "segment capacity exceeded");
        return *h_status;
    }
    if (padded_tail_host) *padded_tail_host = *h_status & 1;
    memcpy(seg_start_host, h_s, sizeof(int64_t) * (*h_cnt));
    memcpy(seg_len_host, h_l, sizeof(int64_t) * (*h_cnt));
    *n_segments_host = *h_cnt;
    return AAT_OK;
}

int aat_host_tokenize(aat_ctx *ctx, const void *wave_host, int wave_dtype, int64_t n_samples, const float *mel_in_host,
                      float *mel_out_host, int64_t *minima_host, int64_t *n_minima_host, int64_t *seg_start_host,
                      int64_t *seg_len_host, int64_t capacity, int64_t *n_segments_host, int32_t *padded_tail_host)
{
    AAT_REQUIRE(ctx && (wave_host || mel_in_host) && seg_len_host && n_segments_host, AAT_ERR_INVALID,
                "aat_host_tokenize: NULL argument");
    AAT_REQUIRE(mel_in_host || wave_dtype == AAT_F32 || wave_dtype == AAT_F64, AAT_ERR_UNSUPPORTED,
                "aat_host_tokenize: waveform dtype must be AAT_F32 or AAT_F64");
    return host_pipeline(ctx, wave_host, wave_dtype, n_samples, mel_in_host, true, mel_out_host, minima_host,
                         n_minima_host, seg_start_host, seg_len_host, capacity, n_segments_host, padded_tail_host);
}

int aat_host_mean_pool(aat_ctx *ctx, const void *emb_host, int emb_dtype, int64_t n_rows, int32_t dim,
                       const int64_t *seg_off_host, int64_t n_seg, float *out_host, double *colsum_host)
{
    AAT_REQUIRE(ctx && (emb_host || n_rows == 0) && seg_off_host && (out_host || n_seg == 0), AAT_ERR_INVALID,
                "aat_host_mean_pool: NULL argument");
    AAT_REQUIRE(n_rows >= 0 && n_seg >= 0 && dim > 0, AAT_ERR_INVALID, "aat_host_mean_pool: negative size");
    const size_t esize = dtype_size(emb_dtype);
    AAT_REQUIRE(emb_dtype == AAT_F32 || emb_dtype == AAT_F16 || emb_dtype == AAT_BF16, AAT_ERR_UNSUPPORTED,
                "aat_host_mean_pool: embedding dtype must be F32, F16 or BF16");
    DeviceGuard guard(ctx->device);
    std::lock_guard<std::mutex> lock(ctx->host_mutex);
    const size_t emb_bytes = esize * (size_t)n_rows * dim, out_bytes = sizeof(float) * (size_t)n_seg * dim;
    const size_t off_bytes = sizeof(int64_t) * ((size_t)n_seg + 1), cs_bytes = sizeof(double) * ((size_t)dim + 1);
    // embeddings are copied straight from the caller's buffer (registering/pinning is the caller's choice)
    const size_t dev_bytes = align256(emb_bytes) + align256(out_bytes) + align256(off_bytes) + align256(cs_bytes);
    int rc = ensure_scratch(ctx, dev_bytes, align256(off_bytes) + align256(cs_bytes));
    if (rc) return rc;
    Arena d(ctx->dev_scratch);
    unsigned char *d_emb = d.take<unsigned char>(emb_bytes);
    float *d_out = d.take<float>((size_t)n_seg * dim);
    int64_t *d_off = d.take<int64_t>((size_t)n_seg + 1);
    double *d_cs = d.take<double>((size_t)dim + 1);
    cudaStream_t st = ctx->host_stream;
    if (emb_bytes) AAT_CUDA_CHECK(cudaMemcpyAsync(d_emb, emb_host, emb_bytes, cudaMemcpyHostToDevice, st));
    AAT_CUDA_CHECK(cudaMemcpyAsync(d_off, seg_off_host, off_bytes, cudaMemcpyHostToDevice, st));
    rc = launch_mean_pool(ctx, nullptr, d_emb, emb_dtype, n_rows, dim, d_off, n_seg, nullptr, d_out,
                          colsum_host ? d_cs : nullptr, 0, st);
    if (rc) return rc;
    if (out_bytes) AAT_CUDA_CHECK(cudaMemcpyAsync(out_host, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    if (colsum_host) AAT_CUDA_CHECK(cudaMemcpyAsync(colsum_host, d_cs, cs_bytes, cudaMemcpyDeviceToHost, st));
    AAT_CUDA_CHECK(cudaStreamSynchronize(st));
    return AAT_OK;
}

// The reference's own calling convention for the pooling (ref:scripts/mean_hubert_embeddings.py:18-20): a list of
// per-segment tensors [1, n_i, D] in host memory.  Every tensor is copied once, straight into the pinned staging
// buffer (no host-side concatenation first), and the DMA of a filled chunk overlaps the CPU copy of the next one.
int aat_host_mean_pool_list(aat_ctx *ctx, const void *const *seg_ptrs_host, const int64_t *seg_rows_host, int64_t n_seg,
                            int emb_dtype, int32_t dim, float *out_host, double *colsum_host)
{
    AAT_REQUIRE(ctx && (n_seg == 0 || (seg_ptrs_host && seg_rows_host && out_host)), AAT_ERR_INVALID,
                "aat_host_mean_pool_list: NULL argument");
    AAT_REQUIRE(n_seg >= 0 && dim > 0, AAT_ERR_INVALID, "aat_host_mean_pool_list: negative size");
    AAT_REQUIRE(emb_dtype == AAT_F32 || emb_dtype == AAT_F16 || emb_dtype == AAT_BF16, AAT_ERR_UNSUPPORTED,
                "aat_host_mean_pool_list: embedding dtype must be F32, F16 or BF16");
    const size_t esize = dtype_size(emb_dtype);
    int64_t n_rows = 0;
    for (int64_t i = 0; i < n_seg; ++i) {
        AAT_REQUIRE(seg_rows_host[i] >= 0 && (seg_rows_host[i] == 0 || seg_ptrs_host[i]), AAT_ERR_INVALID,
                    "aat_host_mean_pool_list: segment %lld has a negative length or a NULL pointer", (long long)i);
        n_rows += seg_rows_host[i];
    }
    DeviceGuard guard(ctx->device);
    std::lock_guard<std::mutex> lock(ctx->host_mutex);
    const size_t row_bytes = esize * (size_t)dim;
    const size_t emb_bytes = row_bytes * (size_t)n_rows, out_bytes = sizeof(float) * (size_t)n_seg * dim;
    const size_t off_bytes = sizeof(int64_t) * ((size_t)n_seg + 1), cs_bytes = sizeof(double) * ((size_t)dim + 1);
    const size_t total = align256(emb_bytes) + align256(out_bytes) + align256(off_bytes) + align256(cs_bytes);
    int rc = ensure_scratch(ctx, total, total);
    if (rc) return rc;
    Arena d(ctx->dev_scratch), h(ctx->pinned);
    unsigned char *d_emb = d.take<unsigned char>(emb_bytes), *h_emb = h.take<unsigned char>(emb_bytes);
    float *d_out = d.take<float>((size_t)n_seg * dim), *h_out = h.take<float>((size_t)n_seg * dim);
    int64_t *d_off = d.take<int64_t>((size_t)n_seg + 1), *h_off = h.take<int64_t>((size_t)n_seg + 1);
    double *d_cs = d.take<double>((size_t)dim + 1), *h_cs = h.take<double>((size_t)dim + 1);
    cudaStream_t st = ctx->host_stream;
    constexpr size_t kChunk = 256 * 1024; // DMA granularity: large enough to run at PCIe speed, small enough to overlap
    size_t filled = 0, sent = 0;
    h_off[0] = 0;
    for (int64_t i = 0; i < n_seg; ++i) {
        const size_t bytes = row_bytes * (size_t)seg_rows_host[i];
        if (bytes) memcpy(h_emb + filled, seg_ptrs_host[i], bytes);
        filled += bytes;
        h_off[i + 1] = h_off[i] + seg_rows_host[i];
        if (filled - sent >= kChunk) {
            AAT_CUDA_CHECK(cudaMemcpyAsync(d_emb + sent, h_emb + sent, filled - sent, cudaMemcpyHostToDevice, st));
            sent = filled;
        }
    }
    if (filled > sent) AAT_CUDA_CHECK(cudaMemcpyAsync(d_emb + sent, h_emb + sent, filled - sent, cudaMemcpyHostToDevice, st));
    AAT_CUDA_CHECK(cudaMemcpyAsync(d_off, h_off, off_bytes, cudaMemcpyHostToDevice, st));
    rc = launch_mean_pool(ctx, nullptr, d_emb, emb_dtype, n_rows, dim, d_off, n_seg, nullptr, d_out,
                          colsum_host ? d_cs : nullptr, 0, st);
    if (rc) return rc;
    if (out_bytes) AAT_CUDA_CHECK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    if (colsum_host) AAT_CUDA_CHECK(cudaMemcpyAsync(h_cs, d_cs, cs_bytes, cudaMemcpyDeviceToHost, st));
    AAT_CUDA_CHECK(cudaStreamSynchronize(st));
    if (out_bytes) memcpy(out_host, h_out, out_bytes);
    if (colsum_host) memcpy(colsum_host, h_cs, cs_bytes);
    return AAT_OK;
}

} // extern "C"
