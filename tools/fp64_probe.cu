// Throughput probe for the K7 design question: can the FP64 pipe carry the 64-bit
// lexicographic compare (DSETP) beside the integer ALU pipe on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/fp64_probe tools/fp64_probe.cu
// Prints lane-ops per clock per SM for: DSETP+2SEL max chains, ISETP pair + 2SEL max chains,
// DADD, and a mix of DSETP-max with VIMNMX3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, UNR = 8;

__device__ __forceinline__ double dmax_bits(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ long long imax64(long long a, long long b) { return a > b ? a : b; }

template <int MODE>
__global__ void probe(double *out, long long *clk, double seed)
{
    double v[UNR], w[UNR];
    long long x[UNR], y[UNR];
    int z[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        v[u] = seed + threadIdx.x + u; w[u] = seed * 1.5 + u;
        x[u] = (long long)(threadIdx.x + u) * 77; y[u] = (long long)u * 1234567 + blockIdx.x;
        z[u] = threadIdx.x + u;
    }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (MODE == 0) { v[u] = dmax_bits(v[u], w[u]); w[u] = dmax_bits(w[u], v[(u + 1) % UNR]); }
            if (MODE == 1) { x[u] = imax64(x[u], y[u]); y[u] = imax64(y[u], x[(u + 1) % UNR]); }
            if (MODE == 2) { v[u] = v[u] + w[u]; w[u] = w[u] + v[(u + 1) % UNR]; }
            if (MODE == 3) {
                v[u] = dmax_bits(v[u], w[u]); w[u] = dmax_bits(w[u], v[(u + 1) % UNR]);
                z[u] = __vimax3_s32(z[u], z[(u + 1) % UNR], it); z[u] = __viaddmax_s32(z[u], 3, z[(u + 2) % UNR]);
            }
            if (MODE == 4) {   // the compare alone: predicate folded into an integer counter
                z[u] += (v[u] > w[(u + it) % UNR]) ? 1 : 0;
            }
        }
    }
    long long t1 = clock64();
    double acc = 0;
#pragma unroll
    for (int u = 0; u < UNR; ++u) acc += v[u] + w[u] + (double)x[u] + (double)y[u] + z[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
double run(int n_sm, int ops_per_iter)
{
    const int grid = n_sm * 4, block = 256;
    double *out; long long *clk;
    cudaMalloc(&out, sizeof(double) * grid * block);
    cudaMalloc(&clk, sizeof(long long) * grid);
    probe<MODE><<<grid, block>>>(out, clk, 1.25);
    probe<MODE><<<grid, block>>>(out, clk, 1.25);
    cudaDeviceSynchronize();
    long long h[4096];
    cudaMemcpy(h, clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += h[i];
    avg /= grid;
    cudaFree(out); cudaFree(clk);
    // 4 CTAs of 256 threads per SM run concurrently
    return 4.0 * block * (double)ITERS * UNR * ops_per_iter / avg;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int n = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d,\n", p.name, n);
    printf(" \"dmax_f64_per_clk_sm\": %.2f,\n", run<0>(n, 2));
    printf(" \"imax_s64_per_clk_sm\": %.2f,\n", run<1>(n, 2));
    printf(" \"dadd_per_clk_sm\": %.2f,\n", run<2>(n, 2));
    printf(" \"dmax_f64_with_2_dpx_per_clk_sm\": %.2f,\n", run<3>(n, 2));
    printf(" \"dsetp_only_per_clk_sm\": %.2f}\n", run<4>(n, 1));
    return 0;
}
