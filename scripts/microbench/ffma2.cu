// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue/pipe throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
constexpr int CH = 8, IT = 4096;
__global__ void k_ffma(float *out, float a, float b)
{
    float x[2 * CH];
    for (int i = 0; i < 2 * CH; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < IT; ++it)
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0;
    for (int i = 0; i < 2 * CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float a, float b)
{
    u64 x[CH];
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    u64 a2 = *reinterpret_cast<u64 *>(&av), b2 = *reinterpret_cast<u64 *>(&bv);
    for (int i = 0; i < CH; ++i) { float2 v = make_float2(threadIdx.x + 2 * i, threadIdx.x + 2 * i + 1); x[i] = *reinterpret_cast<u64 *>(&v); }
    for (int it = 0; it < IT; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fma2(x[i], a2, b2);
    float s = 0;
    for (int i = 0; i < CH; ++i) { float2 v = *reinterpret_cast<float2 *>(&x[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: 1 fma-pipe op + 1 alu-pipe op (LOP3) per element, to see dual issue head-room
__global__ void k_mix(float *out, float a, float b)
{
    float x[CH]; unsigned m[CH];
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x + i; m[i] = i; }
    for (int it = 0; it < IT; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) { x[i] = fmaf(x[i], a, b); m[i] = (m[i] ^ __float_as_uint(a)) + it; }
    float s = 0;
    for (int i = 0; i < CH; ++i) s += x[i] + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_add2(float *out, float a, float b)
{
    u64 x[CH];
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    u64 a2 = *reinterpret_cast<u64 *>(&av), b2 = *reinterpret_cast<u64 *>(&bv);
    for (int i = 0; i < CH; ++i) { float2 v = make_float2(threadIdx.x + 2 * i, threadIdx.x + 2 * i + 1); x[i] = *reinterpret_cast<u64 *>(&v); }
    for (int it = 0; it < IT; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = mul2(add2(x[i], a2), b2);
    float s = 0;
    for (int i = 0; i < CH; ++i) { float2 v = *reinterpret_cast<float2 *>(&x[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K> void run(const char *name, K k, double flop_per_thread, float *out)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * 8, threads = 256;
    k<<<blocks, threads>>>(out, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<<<blocks, threads>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double tf = flop_per_thread * blocks * threads / (ms * 1e-3) / 1e12;
    printf("%-8s %.3f ms  %.1f TFLOP/s (%s)\n", name, ms, tf, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run("ffma", k_ffma, 2.0 * 2 * CH * IT, out);
    run("ffma2", k_ffma2, 2.0 * 2 * CH * IT, out);
    run("mix", k_mix, 2.0 * CH * IT, out);
    run("add2mul2", k_add2, 2.0 * 2 * CH * IT, out);
    return 0;
}
