// l1_gather.cu — what bounds a per-lane gather of 16-/32-byte records through L1 on sm_100a?
// Hypothesis under test: the L1 data stage behaves like shared memory (32 banks x 4 B, LDG.128 served per quarter
// warp), so the cost of a gather is its BANK-CONFLICT degree, not the number of distinct 128-byte lines. Patterns:
//   0 random slot in a 128-slot window                      (what the SPH interaction kernels do today)
//   1 random slot with (slot mod G) == (lane mod G)          (conflict-free inside each quarter warp / LDG.256 phase)
//   2 random slot with (slot mod G) == 0                     (worst case: every lane on the same bank group)
//   3 consecutive slots (lane l reads base + l)              (coalesced reference)
// G = 8 for 16-byte records (8 bank groups of 16 B), 4 for 32-byte records.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o l1_gather l1_gather.cu ; run: ./l1_gather
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int REC> __global__ void __launch_bounds__(128) k_gather(const float4 *__restrict__ rec, const unsigned *__restrict__ idx, int iters, float *out)
{
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned warp = t >> 5, lane = t & 31;
    const unsigned *my = idx + (size_t)warp * iters * 32 + lane;
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < iters; ++k)
    {
        unsigned j = my[32 * k];
        if (REC == 16)
        {
            float4 a = rec[j];
            acc += a.x + a.y * a.z + a.w;
        }
        else
        {
            float4 a, b;
            asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                : "l"(rec + 2ull * j));
            acc += a.x + a.y * a.z + a.w + b.x * b.y + b.z;
        }
    }
    if (acc == 123.456f) out[t] = acc;
}

int main()
{
    const int warps = 148 * 64 * 4, iters = 96, window = 128;
    const size_t nrec = (size_t)warps * 32 + 4096; // each warp's window starts at its own 32 slots (like cell-ordered storage)
    float4 *rec;
    cudaMalloc(&rec, nrec * 32);
    cudaMemset(rec, 0, nrec * 32);
    unsigned *idx;
    cudaMalloc(&idx, (size_t)warps * iters * 32 * 4);
    float *out;
    cudaMalloc(&out, (size_t)warps * 32 * 4);
    std::vector<unsigned> h((size_t)warps * iters * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int recb : {16, 32})
        for (int pat = 0; pat < 4; ++pat)
        {
            const int G = recb == 16 ? 8 : 4;
            srand(1);
            for (int w = 0; w < warps; ++w)
                for (int k = 0; k < iters; ++k)
                    for (int l = 0; l < 32; ++l)
                    {
                        unsigned base = (unsigned)w * 32u, s;
                        if (pat == 0) s = rand() % window;
                        else if (pat == 1) s = (rand() % (window / G)) * G + (l % G);
                        else if (pat == 2) s = (rand() % (window / G)) * G;
                        else s = (k * 7) % (window - 32) + l;
                        h[((size_t)w * iters + k) * 32 + l] = base + s;
                    }
            cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
            for (int rep = 0; rep < 3; ++rep)
            {
                cudaEventRecord(e0);
                if (recb == 16) k_gather<16><<<warps / 4, 128>>>(rec, idx, iters, out);
                else k_gather<32><<<warps / 4, 128>>>(rec, idx, iters, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep == 2)
                    printf("rec %2d B pattern %d: %.3f ms  %.2f warp-gathers/ns  (%.1f SM-cycles per warp gather at 1.9 GHz)\n", recb, pat, ms,
                           (double)warps * iters / (ms * 1e6), ms * 1e-3 * 1.9e9 * 148 / ((double)warps * iters));
            }
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
