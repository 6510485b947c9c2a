// Calibration microbenchmark: issue rate of the DPX / ALU instructions the DP
// kernels are built from, in lane-ops per clock per SM (the
// guides give no DPX rate for sm_100a, so the build measures it).
//
// Each thread runs NCH independent chains x_i = op(x_i, x_{i+1}, a) so the
// compiler can neither fold nor hoist; the grid is 2 CTAs of 512 threads per
// SM (32 warps/SM), every CTA times itself with clock64 and the slowest CTA
// defines the rate.
#include "launchers.h"

namespace bgca {

constexpr int NCH = 8;
constexpr int PROBE_ITERS = 4096;

template <int OP>
__device__ __forceinline__ uint32_t probe_op(uint32_t x, uint32_t y, uint32_t a)
{
    if (OP == 0) return __viaddmax_s16x2(x, a, y);
    if (OP == 1) return __vimax3_s16x2(x, y, a);
    if (OP == 2) return __vmaxs2(x, y);
    if (OP == 3) return __vadd2(x, y);
    if (OP == 4) return x + y + a;
    if (OP == 5) return x * a + y;
    if (OP == 6) return __byte_perm(x, y, a);
    if (OP == 7) return (uint32_t)__viaddmax_s32((int)x, (int)a, (int)y);
    return x;
}

template <int OP>
__global__ void __launch_bounds__(512) probe_kernel(uint32_t a, uint32_t seed, uint32_t *sink,
                                                    long long *cycles)
{
    uint32_t x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < PROBE_ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = probe_op<OP>(x[i], x[(i + 1) % NCH], a);
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// OP 8: the packed local cell recurrence exactly as in gotoh_pair16.cuh (next-T form, gap
// scores as immediates, K = 26 rows per lane, 2 CTAs x 256 threads per SM = the occupancy of the
// real kernel on ~420-residue queries) with all state in registers and NO shared memory,
// shuffles, token handling or loop control beyond the step counter: the ceiling of the DP
// instruction stream itself, which the kernel's remaining ~34 instructions per step eat into.
constexpr int CELL_K = 26;
constexpr int CELL_ITERS = 2048;
__global__ void __launch_bounds__(256, 2) probe_cell_kernel(uint32_t seed, uint32_t *sink, long long *cycles)
{
    constexpr uint32_t EXT2 = 0xFFFFFFFFu, OPEN2 = 0xFFF4FFF4u;   // (-1, -1) and (-12, -12)
    uint32_t Tn[CELL_K], E[CELL_K], P[CELL_K];
#pragma unroll
    for (int k = 0; k < CELL_K; ++k) {
        Tn[k] = 0; E[k] = 0;
        P[k] = ((seed * (threadIdx.x + 3 + k)) & 0x00070007u) + 0x000C000Cu;
    }
    uint32_t best = 0, h_in = seed & 0x000F000Fu, f_in = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < CELL_ITERS; ++it) {
        uint32_t up = h_in, F = f_in, h_prev = 0;
#pragma unroll
        for (int k = 0; k < CELL_K; ++k) {
            F = __viaddmax_s16x2(F, EXT2, up);
            const uint32_t h = __vimax3_s16x2_relu(Tn[k], E[k], F);
            Tn[k] = __vadd2(up, P[k]);
            up = __vadd2(h, OPEN2);
            E[k] = __viaddmax_s16x2(E[k], EXT2, up);
            if (k & 1) best = __vimax3_s16x2(best, h_prev, h);
            h_prev = h;
        }
        h_in = up ^ (uint32_t)(it & 1);   // keep the next step dependent on this one
        f_in = F;
    }
    const long long t1 = clock64();
    uint32_t acc = best ^ h_in ^ f_in;
#pragma unroll
    for (int k = 0; k < CELL_K; ++k) acc ^= Tn[k] ^ E[k];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Mixed-pipe throughput: half of the chains run op A, half op B, independent.
// If A and B issue to different pipes the sum approaches 128 lane-ops/clk/SM.
template <int A, int B>
__global__ void __launch_bounds__(512) probe_mix_kernel(uint32_t a, uint32_t seed, uint32_t *sink,
                                                        long long *cycles)
{
    uint32_t x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < PROBE_ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NCH; i += 2) {
                x[i] = probe_op<A>(x[i], x[(i + 2) % NCH], a);
                x[i + 1] = probe_op<B>(x[i + 1], x[(i + 3) % NCH], a);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Dependent-chain latency: ONE warp per SM, one chain; cycles per op.
// CHAIN 0: viaddmax -> viaddmax ...   1: vimax3 ...   2: vadd2 ...
// CHAIN 3: the cell chain of the DP kernel: F=viaddmax(F,e,Ho); H=vimax3_relu(T,E,F); Ho=vadd2(H,o)
template <int CHAIN>
__global__ void __launch_bounds__(32) probe_latency_kernel(uint32_t a, uint32_t b, uint32_t seed,
                                                           uint32_t *sink, long long *cycles)
{
    uint32_t x = seed * (threadIdx.x + 1), y = seed ^ 0x00050003u, z = seed & 0x00070007u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < PROBE_ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (CHAIN == 0) x = __viaddmax_s16x2(x, a, b);
            if (CHAIN == 1) x = __vimax3_s16x2(x, a, b);
            if (CHAIN == 2) x = __vadd2(x, a);
            if (CHAIN == 3) {
                y = __viaddmax_s16x2(y, a, x);
                const uint32_t h = __vimax3_s16x2_relu(z, b, y);
                x = __vadd2(h, a);
            }
        }
    }
    const long long t1 = clock64();
    if ((x ^ y) == 0x12345678u) sink[0] = x;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CHAIN>
static double run_latency(int n_sm, uint32_t *sink, long long *d_cycles, long long *h_cycles)
{
    probe_latency_kernel<CHAIN><<<n_sm, 32>>>(0xFFFFFFFFu, 0x00010002u, 99u, sink, d_cycles);
    probe_latency_kernel<CHAIN><<<n_sm, 32>>>(0xFFFFFFFFu, 0x00010002u, 99u, sink, d_cycles);
    if (cudaMemcpy(h_cycles, d_cycles, sizeof(long long) * n_sm, cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1.0;
    long long best = h_cycles[0];
    for (int i = 1; i < n_sm; ++i) best = h_cycles[i] < best ? h_cycles[i] : best;
    return (double)best / (8.0 * PROBE_ITERS);   // cycles per op (CHAIN 3: per 3-op cell)
}

template <int A, int B>
static double run_mix(int n_sm, uint32_t *sink, long long *d_cycles, long long *h_cycles)
{
    const int grid = n_sm * 2;
    probe_mix_kernel<A, B><<<grid, 512>>>(0x00030005u, 12345u, sink, d_cycles);
    probe_mix_kernel<A, B><<<grid, 512>>>(0x00030005u, 12345u, sink, d_cycles);
    if (cudaMemcpy(h_cycles, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1.0;
    long long worst = 0;
    for (int i = 0; i < grid; ++i) worst = h_cycles[i] > worst ? h_cycles[i] : worst;
    return 1024.0 * NCH * 4.0 * PROBE_ITERS / (double)worst;
}

template <int OP>
static double run_op(int n_sm, uint32_t *sink, long long *d_cycles, long long *h_cycles)
{
    const int grid = n_sm * 2;
    probe_kernel<OP><<<grid, 512>>>(0x00030005u, 12345u, sink, d_cycles);   // warm-up
    probe_kernel<OP><<<grid, 512>>>(0x00030005u, 12345u, sink, d_cycles);
    if (cudaMemcpy(h_cycles, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1.0;
    long long worst = 0;
    for (int i = 0; i < grid; ++i) worst = h_cycles[i] > worst ? h_cycles[i] : worst;
    const double lane_ops_per_sm = 1024.0 * NCH * 4.0 * PROBE_ITERS;
    return lane_ops_per_sm / (double)worst;
}

int dpx_probe_run(int device, double *out, int n)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    const int n_sm = prop.multiProcessorCount;
    const int grid = n_sm * 2;
    uint32_t *sink = nullptr;
    long long *d_cycles = nullptr;
    if (cudaMalloc(&sink, 64) != cudaSuccess) return -1;
    if (cudaMalloc(&d_cycles, sizeof(long long) * grid) != cudaSuccess) return -1;
    long long *h_cycles = new long long[grid];
    double r[20] = {0};
    r[0] = run_op<0>(n_sm, sink, d_cycles, h_cycles);
    r[1] = run_op<1>(n_sm, sink, d_cycles, h_cycles);
    r[2] = run_op<2>(n_sm, sink, d_cycles, h_cycles);
    r[3] = run_op<3>(n_sm, sink, d_cycles, h_cycles);
    r[4] = run_op<4>(n_sm, sink, d_cycles, h_cycles);
    r[5] = run_op<5>(n_sm, sink, d_cycles, h_cycles);
    r[6] = run_op<6>(n_sm, sink, d_cycles, h_cycles);
    r[7] = run_op<7>(n_sm, sink, d_cycles, h_cycles);
    {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        probe_cell_kernel<<<grid, 256>>>(777u, sink, d_cycles);
        cudaEventRecord(e0);
        probe_cell_kernel<<<grid, 256>>>(777u, sink, d_cycles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h_cycles, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        long long worst = 0;
        for (int i = 0; i < grid; ++i) worst = h_cycles[i] > worst ? h_cycles[i] : worst;
        // lane-ops (5.5 per cell pair) per clock per SM; cell pairs per clock per SM = r[8] / 5.5
        r[8] = 512.0 * (CELL_K * 5.5) * CELL_ITERS / (double)worst;
        r[9] = (double)worst / (ms * 1e3);   // cycles per microsecond = MHz (kernel ~ whole event span)
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    r[10] = run_mix<0, 3>(n_sm, sink, d_cycles, h_cycles);   // VIADDMNMX.S16x2 + VIADD.16x2
    r[11] = run_mix<0, 5>(n_sm, sink, d_cycles, h_cycles);   // VIADDMNMX.S16x2 + IMAD
    r[12] = run_mix<1, 3>(n_sm, sink, d_cycles, h_cycles);   // VIMNMX3.S16x2 + VIADD.16x2
    r[13] = run_mix<3, 5>(n_sm, sink, d_cycles, h_cycles);   // VIADD.16x2 + IMAD
    r[14] = run_latency<0>(n_sm, sink, d_cycles, h_cycles);
    r[15] = run_latency<1>(n_sm, sink, d_cycles, h_cycles);
    r[16] = run_latency<2>(n_sm, sink, d_cycles, h_cycles);
    r[17] = run_latency<3>(n_sm, sink, d_cycles, h_cycles);
    delete[] h_cycles;
    cudaFree(sink);
    cudaFree(d_cycles);
    if (cudaGetLastError() != cudaSuccess) return -1;
    for (int i = 0; i < n && i < 20; ++i) out[i] = r[i];
    return 0;
}

}  // namespace bgca
