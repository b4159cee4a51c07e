// Frame CSR of the WHOLE-UTTERANCE encode convention (synthetic embeddings, convention (ii)).
//
// The pool kernel takes any CSR over embedding rows.  aat_boundaries / aat_segment_frame_csr write the CSR of the
// per-segment encode convention (every segment encoded on its own, ref:scripts/mean_hubert_embeddings.py:18-20).  When
// the encoder ran ONCE over each utterance, utterance b owns T_b = (N_b - 400) / 320 + 1 consecutive rows
// (TF:models/hubert/modeling_hubert.py:675-688, closed form) and a segment that starts at sample s starts at row
// min(s / 320, T_b): the same integer division by the encoder's stride that the collator applies to mel frames
// (ref:src/aat/training/collate.py:340, `// hop_length`).  The sum of the lengths is >= N_b (ref:src/aat/tokenizer.py:195),
// so the last segment of an utterance always ends at T_b and the next utterance starts where it ends.
#include "aat_internal.cuh"

namespace aat {
namespace {

constexpr int kUttCsrThreads = 256;
constexpr int64_t kEncoderField = 400, kEncoderStride = 320; // receptive field and stride of the HuBERT feature encoder

__device__ __forceinline__ int64_t encoder_rows(int64_t n_samples)
{
    return n_samples < kEncoderField ? 0 : (n_samples - kEncoderField) / kEncoderStride + 1;
}

// One CTA per utterance.  The first row of utterance b is the sum of the rows of the utterances before it; every CTA
// adds that up by itself (at most a few thousand cached 8-byte loads), so the kernel needs no scratch and no order
// among its CTAs.
__global__ void __launch_bounds__(kUttCsrThreads)
utterance_frame_csr_kernel(int n_utts, const int64_t *__restrict__ n_samples, const int64_t *__restrict__ seg_slot_off,
                           const int64_t *__restrict__ seg_start, const int32_t *__restrict__ seg_count,
                           const int64_t *__restrict__ utt_seg_off, int64_t *__restrict__ seg_off,
                           int64_t *__restrict__ n_seg_out)
{
    __shared__ int64_t s_part[kUttCsrThreads / 32];
    __shared__ int64_t s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t part = 0;
    for (int i = tid; i < b; i += kUttCsrThreads) part += encoder_rows(n_samples[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (tid == 0) {
        int64_t base = 0;
        for (int w = 0; w < kUttCsrThreads / 32; ++w) base += s_part[w];
        s_base = base;
    }
    __syncthreads();
    const int64_t base = s_base, rows = encoder_rows(n_samples[b]);
    const int cnt = seg_count[b];
    const int64_t first = utt_seg_off[b];
    const int64_t *start = seg_start + seg_slot_off[b];
    for (int j = tid; j < cnt; j += kUttCsrThreads) {
        const int64_t r = start[j] / kEncoderStride;
        seg_off[first + j] = base + (r < rows ? r : rows);
    }
    if (b == n_utts - 1 && tid == 0) {
        seg_off[first + cnt] = base + rows;
        n_seg_out[0] = first + cnt;
        n_seg_out[1] = base + rows;
    }
}

} // namespace

int launch_utterance_frame_csr(const aat_plan *plan, const int64_t *seg_start, const int32_t *seg_count,
                               const int64_t *utt_seg_off, int64_t *seg_off, int64_t *n_seg, cudaStream_t stream)
{
    if (plan->n_utts == 0) { // no utterances: an empty CSR {0}, totals {0, 0}
        AAT_CUDA_CHECK(cudaMemsetAsync(seg_off, 0, sizeof(int64_t), stream));
        AAT_CUDA_CHECK(cudaMemsetAsync(n_seg, 0, 2 * sizeof(int64_t), stream));
        return AAT_OK;
    }
    utterance_frame_csr_kernel<<<plan->n_utts, kUttCsrThreads, 0, stream>>>(
        plan->n_utts, plan->d_n_samples, plan->d_seg_slot_off, seg_start, seg_count, utt_seg_off, seg_off, n_seg);
    AAT_LAUNCH_CHECK();
    return AAT_OK;
}

} // namespace aat
