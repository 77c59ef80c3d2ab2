// tex_gather.cu — does a per-lane gather through the TEXTURE path (tex1Dfetch on linear memory) add to the LSU gather
// throughput on sm_100a, or do both share one L1 data stage? Variants, per warp-row (index load included):
//   0  LDG.128 (16 B)                 1  TEX float4 (16 B)
//   2  LDG.128 + LDG.128 (2 arrays)   3  LDG.128 + TEX float4 (2 arrays)     4  LDG.256 (32-byte record)
// Address patterns: 0 random slot in a 128-slot window, 1 random with (slot mod 8) == (lane mod 8).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tex_gather tex_gather.cu ; run: ./tex_gather
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int V> __global__ void __launch_bounds__(128) k_gather(const float4 *__restrict__ a, const float4 *__restrict__ b, cudaTextureObject_t tb,
                                                                  const unsigned *__restrict__ idx, int iters, float *out)
{
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned warp = t >> 5, lane = t & 31;
    const unsigned *my = idx + (size_t)warp * iters * 32 + lane;
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < iters; ++k)
    {
        unsigned j = my[32 * k];
        if (V == 0 || V == 2 || V == 3)
        {
            float4 x = a[j];
            acc += x.x + x.y * x.z + x.w;
        }
        if (V == 2)
        {
            float4 y = b[j];
            acc += y.x + y.y * y.z + y.w;
        }
        if (V == 1 || V == 3)
        {
            float4 y = tex1Dfetch<float4>(tb, (int)j);
            acc += y.x + y.y * y.z + y.w;
        }
        if (V == 4)
        {
            float4 x, y;
            asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w), "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w)
                : "l"(a + 2ull * j));
            acc += x.x + x.y * x.z + x.w + y.x * y.y + y.z;
        }
    }
    if (acc == 123.456f) out[t] = acc;
}

int main()
{
    const int warps = 148 * 64 * 4, iters = 96, window = 128;
    const size_t nrec = (size_t)warps * 32 + 4096;
    float4 *a, *b;
    cudaMalloc(&a, nrec * 32);
    cudaMalloc(&b, nrec * 16);
    cudaMemset(a, 0, nrec * 32);
    cudaMemset(b, 0, nrec * 16);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = b;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>();
    rd.res.linear.sizeInBytes = nrec * 16;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tb = 0;
    cudaCreateTextureObject(&tb, &rd, &td, nullptr);
    printf("texture object: %s (records %zu)\n", cudaGetErrorString(cudaGetLastError()), nrec);
    unsigned *idx;
    cudaMalloc(&idx, (size_t)warps * iters * 32 * 4);
    float *out;
    cudaMalloc(&out, (size_t)warps * 32 * 4);
    std::vector<unsigned> h((size_t)warps * iters * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[] = {"LDG.128", "TEX float4", "LDG.128 + LDG.128", "LDG.128 + TEX float4", "LDG.256"};
    for (int pat = 0; pat < 2; ++pat)
    {
        srand(1);
        for (int w = 0; w < warps; ++w)
            for (int k = 0; k < iters; ++k)
                for (int l = 0; l < 32; ++l)
                {
                    unsigned base = (unsigned)w * 32u, s;
                    if (pat == 0) s = rand() % window;
                    else s = (rand() % (window / 8)) * 8 + (l % 8);
                    h[((size_t)w * iters + k) * 32 + l] = base + s;
                }
        cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        for (int v = 0; v < 5; ++v)
            for (int rep = 0; rep < 3; ++rep)
            {
                cudaEventRecord(e0);
                switch (v)
                {
                case 0: k_gather<0><<<warps / 4, 128>>>(a, b, tb, idx, iters, out); break;
                case 1: k_gather<1><<<warps / 4, 128>>>(a, b, tb, idx, iters, out); break;
                case 2: k_gather<2><<<warps / 4, 128>>>(a, b, tb, idx, iters, out); break;
                case 3: k_gather<3><<<warps / 4, 128>>>(a, b, tb, idx, iters, out); break;
                default: k_gather<4><<<warps / 4, 128>>>(a, b, tb, idx, iters, out); break;
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep == 2)
                    printf("pattern %d  %-22s %.3f ms  (%.1f SM-cycles per warp row at 1.9 GHz)\n", pat, names[v], ms,
                           ms * 1e-3 * 1.9e9 * 148 / ((double)warps * iters));
            }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
