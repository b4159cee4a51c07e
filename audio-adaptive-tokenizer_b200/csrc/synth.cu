// Synthetic inputs generated ON the device from a counter-based RNG ("16 kHz mono fp32, generated on
// device from a seeded counter-based RNG"), for the benchmark's dataset-scale job (BASELINE config 5: 1000 audio-hours
// of distinct utterances, far more than fits in HBM or crosses PCIe in reasonable time) and for tests.
//
// These kernels produce INPUTS of the path, never results.  The recipe is the one aat_b200/synth.py implements on the
// host with numpy (bursty "syllable" audio: Gaussian noise times an envelope of Hann-shaped voiced bursts of
// U(80, 600) ms and amplitude U(0.3, 1.0), separated by pauses of U(30, 250) ms at a 1e-3 floor), driven by Philox
// 4x32-10 instead of numpy's PCG64: sample i of utterance u is a pure function of (seed + u, i), so any shard of any
// size can be generated independently on any rank.  HBM-write bound (4 B per sample, 4 B per embedding element).
#include "aat_internal.cuh"

namespace aat {

namespace {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const
    {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a, c1 = lo1, c2 = hi0 ^ c3 ^ b, c3 = lo0;
            a += 0x9E3779B9u, b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

__device__ __forceinline__ float u01(uint32_t x) { return ((float)x + 0.5f) * 2.3283064365386963e-10f; } // (0, 1)

// two standard normals from two 32-bit words (Box-Muller)
__device__ __forceinline__ float2 normal_pair(uint32_t x, uint32_t y)
{
    const float r = sqrtf(-2.0f * __logf(u01(x)));
    float s, c;
    sincospif(2.0f * u01(y), &s, &c);
    return make_float2(r * c, r * s);
}

struct alignas(16) Burst {
    long long start; // first sample of the voiced burst
    int len;         // its length in samples (the Hann shape has this many points)
    float amp;
};

constexpr uint32_t kEnvStream = 0x5eed0001u;

// One thread per utterance lays out its bursts: a short sequential process (~2.4 bursts per second of audio).
__global__ void synth_schedule_kernel(int n_utts, const int64_t *n_samples, const int64_t *burst_off, uint64_t seed_base,
                                      int64_t utt_index_base, int sampling_rate, Burst *bursts, int32_t *burst_count)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_utts) return;
    const uint64_t seed = seed_base + (uint64_t)(utt_index_base + b);
    const Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32) ^ kEnvStream};
    const int64_t n = n_samples[b];
    Burst *out = bursts + burst_off[b];
    const int64_t cap = burst_off[b + 1] - burst_off[b];
    const float per_ms = (float)sampling_rate / 1000.0f;
    int64_t pos = 0;
    int k = 0;
    while (pos < n && k < cap) {
        const uint4 r = rng((uint32_t)k, 0u, 0u, 0u);
        const int len = (int)((80.0f + 520.0f * u01(r.x)) * per_ms);
        const float amp = 0.3f + 0.7f * u01(r.y);
        const int pause = (int)((30.0f + 220.0f * u01(r.z)) * per_ms);
        Burst e;
        e.start = pos, e.len = len, e.amp = amp;
        out[k++] = e;
        pos += (int64_t)len + pause;
    }
    burst_count[b] = k;
}

// grid = the plan's 4096-sample chunks; 256 threads, a float4 (= one Philox block) per thread and pass
__global__ void __launch_bounds__(256)
synth_fill_kernel(float *wave, const int64_t *n_samples, const int64_t *wave_off, const int32_t *chunk_utt,
                  const int32_t *chunk_first, const int64_t *burst_off, const Burst *bursts, const int32_t *burst_count,
                  uint64_t seed_base, int64_t utt_index_base)
{
    constexpr int kMaxLocal = 16; // bursts that can touch one 4096-sample chunk: 4096 / (80 + 30 ms) + 2, with slack
    __shared__ Burst s_b[kMaxLocal];
    __shared__ int s_nb;
    const int utt = chunk_utt[blockIdx.x];
    const int64_t c = blockIdx.x - chunk_first[utt];
    const int64_t n = n_samples[utt];
    const int64_t j0 = c * kNormChunk;
    const int64_t len = (n - j0 < kNormChunk) ? n - j0 : kNormChunk;
    if (threadIdx.x == 0) {
        const Burst *tab = bursts + burst_off[utt];
        const int cnt = burst_count[utt];
        int lo = 0, hi = cnt; // bursts [0, lo) start at or before j0
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (tab[mid].start <= j0) lo = mid + 1; else hi = mid;
        }
        int first = lo > 0 ? lo - 1 : 0, m = 0;
        for (int k = first; k < cnt && m < kMaxLocal && tab[k].start < j0 + len; ++k) s_b[m++] = tab[k];
        s_nb = m;
    }
    __syncthreads();
    const uint64_t seed = seed_base + (uint64_t)(utt_index_base + utt);
    const Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32)};
    float *dst = wave + wave_off[utt] + j0;
    const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    const int nb = s_nb;
    for (int i = 4 * threadIdx.x; i < len; i += 4 * 256) {
        const int64_t g = j0 + i; // multiple of 4: Philox block g / 4 of this utterance
        const uint4 r = rng((uint32_t)(g >> 2), (uint32_t)((uint64_t)g >> 34), 0u, 0u);
        const float2 z0 = normal_pair(r.x, r.y), z1 = normal_pair(r.z, r.w);
        float v[4] = {z0.x, z0.y, z1.x, z1.y};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t s = g + u;
            float env = 1e-3f;
            for (int k = 0; k < nb; ++k) {
                const int64_t d = s - s_b[k].start;
                if (d >= 0 && d < s_b[k].len) { // np.hanning(len)[d] * amp + floor
                    const float w = s_b[k].len > 1 ? 0.5f - 0.5f * cospif(2.0f * (float)d / (float)(s_b[k].len - 1)) : 1.0f;
                    env = w * s_b[k].amp + 1e-3f;
                }
            }
            v[u] *= env;
        }
        if (vec && i + 4 <= len) {
            *reinterpret_cast<float4 *>(dst + i) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u < len) dst[i + u] = v[u];
        }
    }
}

// out[i] ~ N(0, 1): element i is a pure function of (seed, i).  Grid-stride, a float4 per thread and pass.
__global__ void __launch_bounds__(256) synth_normal_kernel(float *out, int64_t n, uint64_t seed)
{
    const Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x0e3bedd5u};
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        const uint4 r = rng((uint32_t)q, (uint32_t)((uint64_t)q >> 32), 0u, 0u);
        const float2 z0 = normal_pair(r.x, r.y), z1 = normal_pair(r.z, r.w);
        *reinterpret_cast<float4 *>(out + 4 * q) = make_float4(z0.x, z0.y, z1.x, z1.y);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { // the last 1-3 elements
        const int64_t q = n4;
        const uint4 r = rng((uint32_t)q, (uint32_t)((uint64_t)q >> 32), 0u, 0u);
        const float2 z0 = normal_pair(r.x, r.y), z1 = normal_pair(r.z, r.w);
        const float v[4] = {z0.x, z0.y, z1.x, z1.y};
        out[4 * q + threadIdx.x] = v[threadIdx.x];
    }
}

} // namespace

// bursts an utterance of n samples can hold: every burst + pause advances by at least (80 + 30) ms
int64_t synth_burst_capacity(int sampling_rate, int64_t n)
{
    const int64_t min_period = (int64_t)(80.0f * sampling_rate / 1000.0f) + (int64_t)(30.0f * sampling_rate / 1000.0f);
    return n / (min_period > 0 ? min_period : 1) + 2;
}

int launch_synth_waveforms(aat_ctx *ctx, const aat_plan *plan, uint64_t seed_base, int64_t utt_index_base, float *wave,
                           void *workspace, cudaStream_t stream)
{
    if (plan->n_utts == 0) return AAT_OK;
    AAT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, AAT_ERR_INVALID,
                "aat_synth_waveforms: workspace must be 16-byte aligned");
    Burst *bursts = static_cast<Burst *>(workspace);
    int32_t *counts = reinterpret_cast<int32_t *>(bursts + plan->total_bursts);
    synth_schedule_kernel<<<(plan->n_utts + 63) / 64, 64, 0, stream>>>(plan->n_utts, plan->d_n_samples, plan->d_burst_off,
                                                                      seed_base, utt_index_base, ctx->cfg.sampling_rate,
                                                                      bursts, counts);
    AAT_LAUNCH_CHECK();
    synth_fill_kernel<<<plan->norm_chunks, 256, 0, stream>>>(wave, plan->d_n_samples, plan->d_wave_off, plan->d_chunk_utt,
                                                            plan->d_chunk_first, plan->d_burst_off, bursts, counts,
                                                            seed_base, utt_index_base);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

size_t synth_workspace_bytes(const aat_plan *plan)
{
    return sizeof(Burst) * (size_t)plan->total_bursts + sizeof(int32_t) * (size_t)(plan->n_utts ? plan->n_utts : 1);
}

int launch_synth_normal(aat_ctx *ctx, float *out, int64_t n, uint64_t seed, cudaStream_t stream)
{
    if (n <= 0) return AAT_OK;
    AAT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, AAT_ERR_INVALID, "aat_synth_normal: out_dev must be 16-byte aligned");
    const int64_t blocks_needed = ((n >> 2) + 255) / 256;
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    const int grid = (int)(blocks_needed < 1 ? 1 : (blocks_needed < cap ? blocks_needed : cap));
    synth_normal_kernel<<<grid, 256, 0, stream>>>(out, n, seed);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat
