// Pure read-stream micro-benchmark of the pool kernel's mechanism, three questions:
//   * how fast can this grid / ring / consumer arrangement stream at all (the ceiling the pool kernel is judged by);
//   * does dynamic work distribution win back the uneven finish of the CTAs (variant C below);
//   * A/B of the two ways the TMA engine can feed the ring:
//   (A) 1-D bulk copies  cp.async.bulk.shared::cluster.global            (what csrc/pool.cu ships: SASS UBLKCP)
//   (B) 2-D tensor-map tile loads  cp.async.bulk.tensor.2d ... .tile     (SASS UTMALDG)
// Same persistent grid (2 CTAs per SM), same 4 x 24 KB ring, same consumers (one 16-byte column slab per thread, four
// interleaved float32 accumulators per column), same bytes: a [rows, 768] float32 matrix streamed once.  The rows of
// the embedding matrix are contiguous and full-width, so a stage of 8 rows is ONE contiguous 24 KB range for (A); for
// (B) the box limit of 256 elements per dimension splits it into three {256 x 8} tiles per stage.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/_build/ubench_tma2d profiles/ubench_tma2d.cu
//   profiles/_build/ubench_tma2d > profiles/r2_ubench_tma2d.txt          (on a B200)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (x);                                                                  \
        if (e__ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int kDim = 768, kStages = 4, kRowsPerStage = 8, kConsumers = kDim / 4, kThreads = kConsumers + 32;
constexpr int kStageBytes = kRowsPerStage * kDim * 4; // 24 KB
constexpr int kBoxCols = 256;                         // tensor-map box limit per dimension

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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tensor2d_g2s(void *dst, const CUtensorMap *map, int col, int row, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(col), "r"(row), "r"(smem_u32(bar))
                 : "memory");
}

// kTensor = false: stage layout [8 rows][768]; true: [3 column blocks][8 rows][256]
template <bool kTensor>
__global__ void __launch_bounds__(kThreads) stream_kernel(const float *emb, const __grid_constant__ CUtensorMap map, int64_t n_rows,
                                                          float *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
    const int tid = threadIdx.x;
    const int64_t G = gridDim.x, c = blockIdx.x;
    const int64_t r0 = (c * n_rows) / G / kRowsPerStage * kRowsPerStage, r1 = ((c + 1) * n_rows) / G / kRowsPerStage * kRowsPerStage;
    const int64_t n_chunks = (r1 - r0) / kRowsPerStage;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1), mbar_init(&s_empty[s], kConsumers / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == kConsumers) { // producer
        for (int64_t ch = 0; ch < n_chunks; ++ch) {
            const int s = (int)(ch % kStages);
            if (ch >= kStages) mbar_wait(&s_empty[s], (uint32_t)(((ch / kStages) - 1) & 1));
            const int64_t row = r0 + ch * kRowsPerStage;
            mbar_expect_tx(&s_full[s], kStageBytes);
            if (kTensor) {
                for (int b = 0; b < kDim / kBoxCols; ++b)
                    tensor2d_g2s(smem + s * kStageBytes + b * (kStageBytes / (kDim / kBoxCols)), &map, b * kBoxCols, (int)row,
                                 &s_full[s]);
            } else {
                bulk_g2s(smem + s * kStageBytes, emb + row * kDim, kStageBytes, &s_full[s]);
            }
        }
        return;
    }
    if (tid > kConsumers) return;
    float acc[4][4] = {};
    // column slab of this thread inside a stage
    const int blk = (tid * 4) / kBoxCols, col_in_blk = (tid * 4) % kBoxCols;
    for (int64_t ch = 0; ch < n_chunks; ++ch) {
        const int s = (int)(ch % kStages);
        mbar_wait(&s_full[s], (uint32_t)((ch / kStages) & 1));
        const float *stage = reinterpret_cast<const float *>(smem + s * kStageBytes);
#pragma unroll
        for (int r = 0; r < kRowsPerStage; ++r) {
            const float *p = kTensor ? stage + (blk * kRowsPerStage + r) * kBoxCols + col_in_blk : stage + r * kDim + tid * 4;
            const float4 x = *reinterpret_cast<const float4 *>(p);
            acc[r & 3][0] += x.x, acc[r & 3][1] += x.y, acc[r & 3][2] += x.z, acc[r & 3][3] += x.w;
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&s_empty[s]);
    }
    float4 o;
    o.x = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
    o.y = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    o.z = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
    o.w = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
    *reinterpret_cast<float4 *>(out + c * kDim + tid * 4) = o;
}

// (C) dynamic balancing: every CTA first streams a static share (kStaticPct % of a fair share), then takes chunks of
// kDynStages stages from a global counter until the rows are used up — how much of the pool kernel's uneven finish
// (profiles/r1_pool_timeline.txt) dynamic work distribution could win back.
__device__ unsigned int g_next_chunk;
template <int kStaticPct, int kDynStages>
__global__ void __launch_bounds__(kThreads) stream_dynamic_kernel(const float *emb, int64_t n_rows, float *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
    __shared__ int64_t s_row[kStages]; // first row of the stage (-1: no more work)
    const int tid = threadIdx.x;
    const int64_t G = gridDim.x, c = blockIdx.x;
    const int64_t stages_total = n_rows / kRowsPerStage;
    const int64_t static_stages = (stages_total * kStaticPct / 100) / G; // per CTA
    const int64_t dyn_begin = static_stages * G;                          // first dynamic stage
    const int64_t n_dyn_chunks = (stages_total - dyn_begin + kDynStages - 1) / kDynStages;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1), mbar_init(&s_empty[s], kConsumers / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == kConsumers) { // producer: static share, then chunks from the counter; a stage with row -1 ends the stream
        int64_t issued = 0;
        auto issue = [&](int64_t stage_idx) {
            const int s = (int)(issued % kStages);
            if (issued >= kStages) mbar_wait(&s_empty[s], (uint32_t)(((issued / kStages) - 1) & 1));
            s_row[s] = stage_idx;
            if (stage_idx >= 0) {
                mbar_expect_tx(&s_full[s], kStageBytes);
                bulk_g2s(smem + s * kStageBytes, emb + stage_idx * kRowsPerStage * kDim, kStageBytes, &s_full[s]);
            } else {
                mbar_arrive(&s_full[s]);
            }
            ++issued;
        };
        for (int64_t i = 0; i < static_stages; ++i) issue(c * static_stages + i);
        for (;;) {
            const unsigned int k = atomicAdd(&g_next_chunk, 1u);
            if ((int64_t)k >= n_dyn_chunks) break;
            for (int j = 0; j < kDynStages; ++j) {
                const int64_t st = dyn_begin + (int64_t)k * kDynStages + j;
                if (st < stages_total) issue(st);
            }
        }
        issue(-1);
        return;
    }
    if (tid > kConsumers) return;
    float acc[4][4] = {};
    for (int64_t ch = 0;; ++ch) {
        const int s = (int)(ch % kStages);
        mbar_wait(&s_full[s], (uint32_t)((ch / kStages) & 1));
        if (s_row[s] < 0) break;
        const float *stage = reinterpret_cast<const float *>(smem + s * kStageBytes);
#pragma unroll
        for (int r = 0; r < kRowsPerStage; ++r) {
            const float4 x = *reinterpret_cast<const float4 *>(stage + r * kDim + tid * 4);
            acc[r & 3][0] += x.x, acc[r & 3][1] += x.y, acc[r & 3][2] += x.z, acc[r & 3][3] += x.w;
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&s_empty[s]);
    }
    float4 o;
    o.x = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
    o.y = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    o.z = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
    o.w = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
    *reinterpret_cast<float4 *>(out + c * kDim + tid * 4) = o;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    CK(cudaSetDevice(0));
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, 0));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        fprintf(stderr, "cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    const int grid = prop.multiProcessorCount * 2;
    printf("# %s, %d SMs; grid %d x %d threads, ring %d x %d KB; float32 [rows, %d] streamed once, 4 rotating inputs,\n", prop.name,
           prop.multiProcessorCount, grid, kThreads, kStages, kStageBytes / 1024, kDim);
    printf("# 32 launches back to back inside one event pair, best of 5\n");
    printf("%10s %10s %22s %22s %28s %28s\n", "rows", "MB", "1-D bulk: us   GB/s", "2-D tensor map: us   GB/s",
           "75% static + dyn(2): us GB/s", "50% static + dyn(2): us GB/s");
    float *out = nullptr;
    CK(cudaMalloc(&out, sizeof(float) * (size_t)grid * kDim));
    CK(cudaFuncSetAttribute(stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes));
    CK(cudaFuncSetAttribute(stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes));
    CK(cudaFuncSetAttribute(stream_dynamic_kernel<75, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes));
    CK(cudaFuncSetAttribute(stream_dynamic_kernel<50, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes));
    for (int64_t n_rows : {49568LL, 255744LL, 719872LL}) {
        const int R = n_rows * kDim * 4 > 1500000000LL ? 2 : 4;
        std::vector<float *> emb(R);
        std::vector<CUtensorMap> maps(R);
        for (int i = 0; i < R; ++i) {
            CK(cudaMalloc(&emb[i], sizeof(float) * (size_t)n_rows * kDim));
            CK(cudaMemset(emb[i], 0, sizeof(float) * (size_t)n_rows * kDim));
            const cuuint64_t dims[2] = {(cuuint64_t)kDim, (cuuint64_t)n_rows};
            const cuuint64_t strides[1] = {(cuuint64_t)kDim * 4};
            const cuuint32_t box[2] = {kBoxCols, kRowsPerStage}, estr[2] = {1, 1};
            const CUresult r = ((EncodeFn)fn)(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, emb[i], dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
                return 1;
            }
        }
        double us[4];
        for (int variant = 0; variant < 4; ++variant) {
            auto launch = [&](int i) {
                if (variant >= 2) {
                    unsigned int zero = 0; // the counter is reset by a tiny async copy in front of every launch
                    CK(cudaMemcpyToSymbolAsync(g_next_chunk, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice));
                }
                if (variant == 0)
                    stream_kernel<false><<<grid, kThreads, kStages * kStageBytes>>>(emb[i % R], maps[i % R], n_rows, out);
                else if (variant == 1)
                    stream_kernel<true><<<grid, kThreads, kStages * kStageBytes>>>(emb[i % R], maps[i % R], n_rows, out);
                else if (variant == 2)
                    stream_dynamic_kernel<75, 2><<<grid, kThreads, kStages * kStageBytes>>>(emb[i % R], n_rows, out);
                else
                    stream_dynamic_kernel<50, 2><<<grid, kThreads, kStages * kStageBytes>>>(emb[i % R], n_rows, out);
            };
            for (int i = 0; i < 8; ++i) launch(i);
            CK(cudaDeviceSynchronize());
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaEventRecord(e0));
                for (int i = 0; i < 32; ++i) launch(i);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            us[variant] = best * 1e3 / 32;
        }
        const double mb = (double)n_rows * kDim * 4 / 1e6;
        printf("%10lld %10.1f %12.2f %9.0f %14.2f %9.0f %18.2f %9.0f %18.2f %9.0f\n", (long long)n_rows, mb, us[0], mb / us[0] * 1e3, us[1],
               mb / us[1] * 1e3, us[2], mb / us[2] * 1e3, us[3], mb / us[3] * 1e3);
        for (int i = 0; i < R; ++i) CK(cudaFree(emb[i]));
    }
    return 0;
}
