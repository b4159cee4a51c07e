// Micro-benchmark behind the log-mel kernel's design choices: issue cost and dependent latency of the FP64
// and conversion instructions on one B200 SM.   nvcc -arch=sm_100a -O3 -o /tmp/ub profiles/ubench_fp64.cu && /tmp/ub
// throughput: 16 warps (4 per scheduler) x 8 independent chains; latency: 1 warp x 1 chain; cycles from clock64().
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

enum Op { DFMA, DADD, DMUL, F2D, D2F, I2D, D2F2D, FFMA, FADD32 };

template <int OP>
__device__ __forceinline__ double step(double x, double a, double b)
{
    if (OP == DFMA) return fma(x, a, b);
    if (OP == DADD) return x + b;
    if (OP == DMUL) return x * a;
    if (OP == F2D) return (double)(__double_as_longlong(x) == 7 ? 0.f : __int_as_float((int)(__double_as_longlong(x) >> 32))); // F2F.F64.F32 on the high word
    if (OP == D2F) return __hiloint2double(__float_as_int((float)x), 0x12345678);                                               // F2F.F32.F64
    if (OP == I2D) return (double)(int)(__double_as_longlong(x) >> 40);                                                        // I2F.F64.S32
    if (OP == D2F2D) return (double)(float)x;
    if (OP == FADD32) { // float chain carried in the low word: one FADD per step
        const float f = __fadd_rn(__int_as_float(__double2loint(x)), (float)b);
        return __hiloint2double(__double2hiint(x), __float_as_int(f));
    }
    return x;
}

template <int OP, int CHAINS>
__global__ void bench(double *out, long long *cycles, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = 1.0 + threadIdx.x * 1e-3 + c;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = step<OP>(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, double *out, long long *cyc)
{
    long long h;
    bench<OP, 1><<<1, 32>>>(out, cyc, 1.0000001, 1e-9);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double lat = (double)h / (kIters * 4);
    bench<OP, 8><<<1, 512>>>(out, cyc, 1.0000001, 1e-9);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // 16 warps x 8 chains x 4 x kIters warp-instructions on 4 schedulers
    const double per_sched = (double)h / (4.0 * 8 * 4 * kIters);
    printf("%-28s dependent latency %6.1f cycles   issue cost %5.2f cycles per warp-instruction per scheduler\n", name, lat,
           per_sched);
}

int main()
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(double) * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 8);
    run<DFMA>("DFMA", out, cyc);
    run<DADD>("DADD", out, cyc);
    run<DMUL>("DMUL", out, cyc);
    run<F2D>("F2F.F64.F32 (+shift)", out, cyc);
    run<D2F>("F2F.F32.F64 (+mov)", out, cyc);
    run<I2D>("I2F.F64.S32 (+shift)", out, cyc);
    run<D2F2D>("double(float(x)) round trip", out, cyc);
    run<FADD32>("FADD (float32, _rn)", out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
