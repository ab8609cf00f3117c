// Latency and per-SM throughput of the instructions the K5 bound kernel and the Gotoh kernel are built from, on the GPU
// at hand: dependent chains (1 warp) for latency, 32 warps x 8 independent chains for throughput.
// nvcc -gencode arch=compute_100a,code=sm_100a -o dpx_rates dpx_rates.cu && ./dpx_rates
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned s) { unsigned d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }

template <int OP> __device__ __forceinline__ unsigned op(unsigned x, unsigned y, unsigned z, unsigned one)
{
    if (OP == 0) return __vimax3_s16x2(x, y, z);
    if (OP == 1) return (unsigned)__vimax3_s32((int)x, (int)y, (int)z);
    if (OP == 2) return __viaddmax_s16x2(x, y, z);
    if (OP == 3) return (unsigned)__viaddmax_s32((int)x, (int)y, (int)z);
    if (OP == 4) return prmt(x, y, z);
    if (OP == 5) return x * one + y;                 // IMAD
    if (OP == 6) return (x & y) | z;                 // LOP3
    if (OP == 7) return __vimax3_s16x2(x * one + y, y, z);   // IMAD -> VIMNMX3.S16x2 (one K5 row)
    return x;
}

template <int OP, int CHAINS>
__global__ void bench(unsigned *out, long long *cycles, int iters, unsigned one, unsigned seed)
{
    unsigned v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = seed + threadIdx.x * 17 + c;
    const unsigned y = seed * 3 + 1, z = seed ^ 0x00050005u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) v[c] = op<OP>(v[c], y, z, one);
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s ^= v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, unsigned *out, long long *cyc)
{
    const int iters = 2000;
    long long h = 0;
    bench<OP, 1><<<1, 32>>>(out, cyc, iters, 1u, 12345u);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double lat = (double)h / (iters * 8.0);
    bench<OP, 8><<<1, 1024>>>(out, cyc, iters, 1u, 12345u);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double thr = (double)iters * 8 * 8 * 1024 / (double)h;       // lane-ops per clock per SM
    printf("%-34s latency %6.2f cycles   throughput %6.1f lane-ops/clk/SM\n", name, lat, thr);
}

int main()
{
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 64);
    run<0>("VIMNMX3.S16x2", out, cyc);
    run<1>("VIMNMX3.S32", out, cyc);
    run<2>("VIADDMNMX.S16x2", out, cyc);
    run<3>("VIADDMNMX.S32", out, cyc);
    run<4>("PRMT", out, cyc);
    run<5>("IMAD", out, cyc);
    run<6>("LOP3", out, cyc);
    run<7>("IMAD -> VIMNMX3.S16x2 (2 ops)", out, cyc);
    return cudaDeviceSynchronize() != cudaSuccess;
}
