// Stand-alone microbenchmark of candidate inner-loop formulations of the packed local cell
// (tuning aid, not part of the library):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 \
//   -o tools/_build/cell_probe tools/cell_probe.cu && tools/_build/cell_probe
// Each variant sweeps K rows per step with all state in registers, 2 CTAs x 256 threads per SM
// (the occupancy of the real kernel at K = 26), and reports cell-pairs per clock per SM and the
// equivalent GCUPS at the measured clock.  No shared memory, shuffles or token handling: this is
// the ceiling of the DP instruction stream itself.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int K = 26;
constexpr int ITERS = 2048;
constexpr uint32_t EXT2 = 0xFFFFFFFFu, OPEN2 = 0xFFF4FFF4u;

__device__ __forceinline__ uint32_t hmax2_bits(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// VARIANT 0: v4 recurrence as in gotoh_pair16.cuh (best = max3 every 2nd cell)
// VARIANT 1: no best tracking at all (upper bound)
// VARIANT 2: best via max.f16x2 on the raw bits, one per cell
// VARIANT 3: best via max.f16x2, pairwise tree (hmax2(h_prev,h) then hmax2(best, .)) = 1 per cell
// VARIANT 4: best via max3 every 2nd cell but into 2 independent accumulators
template <int VARIANT>
__global__ void __launch_bounds__(256, 2) cell_kernel(uint32_t seed, uint32_t *sink, long long *cycles)
{
    uint32_t Tn[K], E[K], P[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        Tn[k] = 0; E[k] = 0;
        P[k] = ((seed * (threadIdx.x + 3 + k)) & 0x00070007u) + 0x000C000Cu;
    }
    uint32_t best = 0, best2 = 0, h_in = seed & 0x000F000Fu, f_in = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        uint32_t up = h_in, F = f_in, h_prev = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            F = __viaddmax_s16x2(F, EXT2, up);
            const uint32_t h = __vimax3_s16x2_relu(Tn[k], E[k], F);
            Tn[k] = __vadd2(up, P[k]);
            up = __vadd2(h, OPEN2);
            E[k] = __viaddmax_s16x2(E[k], EXT2, up);
            if (VARIANT == 0) {
                if (k & 1) best = __vimax3_s16x2(best, h_prev, h);
            } else if (VARIANT == 2) {
                best = hmax2_bits(best, h);
            } else if (VARIANT == 3) {
                if (k & 1) best = hmax2_bits(best, hmax2_bits(h_prev, h));
            } else if (VARIANT == 4) {
                if ((k & 3) == 1) best = __vimax3_s16x2(best, h_prev, h);
                if ((k & 3) == 3) best2 = __vimax3_s16x2(best2, h_prev, h);
            }
            h_prev = h;
        }
        h_in = up ^ (uint32_t)(it & 1);
        f_in = F;
    }
    const long long t1 = clock64();
    uint32_t acc = best ^ best2 ^ h_in ^ f_in;
#pragma unroll
    for (int k = 0; k < K; ++k) acc ^= Tn[k] ^ E[k];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// VARIANT 5/6: the K rows split in two halves that run one column apart (the lower half
// continues, in this step, the column the upper half finished in the previous step), so every
// step carries TWO independent F->H->Ho chains.  6 = same with K split in three.
template <int PARTS>
__global__ void __launch_bounds__(256, 2) cell_split_kernel(uint32_t seed, uint32_t *sink, long long *cycles)
{
    uint32_t Tn[K], E[K], P[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        Tn[k] = 0; E[k] = 0;
        P[k] = ((seed * (threadIdx.x + 3 + k)) & 0x00070007u) + 0x000C000Cu;
    }
    constexpr int SEG = (K + PARTS - 1) / PARTS;
    uint32_t best = 0, up_in[PARTS], f_in[PARTS];
#pragma unroll
    for (int p = 0; p < PARTS; ++p) { up_in[p] = seed & 0x000F000Fu; f_in[p] = 0; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        uint32_t up[PARTS], F[PARTS], hp[PARTS];
#pragma unroll
        for (int p = 0; p < PARTS; ++p) { up[p] = up_in[p]; F[p] = f_in[p]; hp[p] = 0; }
#pragma unroll
        for (int r = 0; r < SEG; ++r) {
#pragma unroll
            for (int p = 0; p < PARTS; ++p) {
                const int k = p * SEG + r;
                if (k < K) {
                    F[p] = __viaddmax_s16x2(F[p], EXT2, up[p]);
                    const uint32_t h = __vimax3_s16x2_relu(Tn[k], E[k], F[p]);
                    Tn[k] = __vadd2(up[p], P[k]);
                    up[p] = __vadd2(h, OPEN2);
                    E[k] = __viaddmax_s16x2(E[k], EXT2, up[p]);
                    if (r & 1) best = __vimax3_s16x2(best, hp[p], h);
                    hp[p] = h;
                }
            }
        }
        // part p+1 continues next step what part p produced in this one
#pragma unroll
        for (int p = PARTS - 1; p > 0; --p) { up_in[p] = up[p - 1]; f_in[p] = F[p - 1]; }
        up_in[0] = up[PARTS - 1] ^ (uint32_t)(it & 1);
        f_in[0] = F[PARTS - 1];
    }
    const long long t1 = clock64();
    uint32_t acc = best;
#pragma unroll
    for (int p = 0; p < PARTS; ++p) acc ^= up_in[p] ^ f_in[p];
#pragma unroll
    for (int k = 0; k < K; ++k) acc ^= Tn[k] ^ E[k];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// single-op and two-op pipe tests for max.f16x2
template <int MODE>
__global__ void __launch_bounds__(512) pipe_kernel(uint32_t a, uint32_t seed, uint32_t *sink, long long *cycles)
{
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (seed * (threadIdx.x + 1) + i * 0x9E3779B9u) & 0x3FFF3FFFu;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = hmax2_bits(x[i], x[(i + 1) % 8]);
                if (MODE == 1) x[i] = (i & 1) ? hmax2_bits(x[i], x[(i + 2) % 8]) : __vimax3_s16x2(x[i], x[(i + 2) % 8], a);
                if (MODE == 2) x[i] = (i & 1) ? hmax2_bits(x[i], x[(i + 2) % 8]) : __vadd2(x[i], x[(i + 2) % 8]);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Independent-chain mixes at the DP kernel's op ratio: per group of 11 ops, 7 alu (4 viaddmax,
// 3 vimax3) and 4 fma-pipe packed adds, no dependence between neighbouring ops.
// MODE 0: that 7:4 mix; 1: alu-only part (7 ops); 2: 1:1 viaddmax+vimax3; 3: 3 alu : 1 add
template <int MODE>
__global__ void __launch_bounds__(256, 2) ratio_kernel(uint32_t a, uint32_t seed, uint32_t *sink, long long *cycles)
{
    constexpr int N = 22;
    uint32_t x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = (seed * (threadIdx.x + 1) + i * 0x9E3779B9u) & 0x3FFF3FFFu;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int j = i % 11;
                const uint32_t y = x[(i + 5) % N], z = x[(i + 12) % N];
                if (MODE == 0) {
                    if (j == 1 || j == 4 || j == 7 || j == 10) x[i] = __vadd2(x[i], y);
                    else if (j == 2 || j == 5 || j == 8) x[i] = __vimax3_s16x2_relu(x[i], y, z);
                    else x[i] = __viaddmax_s16x2(x[i], 0xFFFFFFFFu, y);
                } else if (MODE == 1) {
                    if (j == 1 || j == 4 || j == 7 || j == 10) continue;
                    else if (j == 2 || j == 5 || j == 8) x[i] = __vimax3_s16x2_relu(x[i], y, z);
                    else x[i] = __viaddmax_s16x2(x[i], 0xFFFFFFFFu, y);
                } else if (MODE == 2) {
                    if (i & 1) x[i] = __vimax3_s16x2_relu(x[i], y, z);
                    else x[i] = __viaddmax_s16x2(x[i], 0xFFFFFFFFu, y);
                } else {
                    if ((i & 3) == 3) x[i] = __vadd2(x[i], y);
                    else if (i & 1) x[i] = __vimax3_s16x2_relu(x[i], y, z);
                    else x[i] = __viaddmax_s16x2(x[i], 0xFFFFFFFFu, y);
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
static double run(F launch, int grid, long long *d_cycles, double *mhz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long worst = 0;
    for (auto c : h) worst = c > worst ? c : worst;
    *mhz = (double)worst / (ms * 1e3);
    return (double)worst;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int n_sm = prop.multiProcessorCount;
    uint32_t *sink; long long *d_cycles;
    cudaMalloc(&sink, 64);
    cudaMalloc(&d_cycles, sizeof(long long) * n_sm * 2);
    double mhz = 0;
    const char *names[] = {"v4 (max3 best every 2nd cell)", "no best (upper bound)", "best = max.f16x2 per cell",
                           "best = max.f16x2 tree", "two max3 accumulators"};
    printf("{\"sm\": %d", n_sm);
#define RUN_CELL(V)                                                                                   \
    {                                                                                                 \
        const int grid = n_sm * 2;                                                                    \
        double cyc = run([&] { cell_kernel<V><<<grid, 256>>>(777u, sink, d_cycles); }, grid, d_cycles, &mhz); \
        double pairs_per_clk = 512.0 * K * ITERS / cyc;                                               \
        printf(", \"cell_v%d\": {\"what\": \"%s\", \"cell_pairs_per_clk_sm\": %.3f, \"gcups_at_1965\": %.1f, \"mhz\": %.0f}", \
               V, names[V], pairs_per_clk, pairs_per_clk * 2 * n_sm * 1.965, mhz);                   \
    }
    RUN_CELL(0) RUN_CELL(1) RUN_CELL(2) RUN_CELL(3) RUN_CELL(4)
#define RUN_SPLIT(PARTS)                                                                              \
    {                                                                                                 \
        const int grid = n_sm * 2;                                                                    \
        double cyc = run([&] { cell_split_kernel<PARTS><<<grid, 256>>>(777u, sink, d_cycles); }, grid, d_cycles, &mhz); \
        double pairs_per_clk = 512.0 * K * ITERS / cyc;                                               \
        printf(", \"cell_split%d\": {\"cell_pairs_per_clk_sm\": %.3f, \"gcups_at_1965\": %.1f}", PARTS,  \
               pairs_per_clk, pairs_per_clk * 2 * n_sm * 1.965);                                      \
    }
    RUN_SPLIT(1) RUN_SPLIT(2) RUN_SPLIT(3)
#define RUN_PIPE(M, label)                                                                            \
    {                                                                                                 \
        const int grid = n_sm * 2;                                                                    \
        double cyc = run([&] { pipe_kernel<M><<<grid, 512>>>(0x00030005u, 12345u, sink, d_cycles); }, grid, d_cycles, &mhz); \
        printf(", \"%s\": %.2f", label, 1024.0 * 8 * 4 * 4096 / cyc);                                 \
    }
    RUN_PIPE(0, "hmnmx2_lane_ops_per_clk_sm") RUN_PIPE(1, "mix_hmnmx2+vimnmx3") RUN_PIPE(2, "mix_hmnmx2+viadd16x2")
#define RUN_RATIO(M, label, nops)                                                                     \
    {                                                                                                 \
        const int grid = n_sm * 2;                                                                    \
        double cyc = run([&] { ratio_kernel<M><<<grid, 256>>>(0x00030005u, 12345u, sink, d_cycles); }, grid, d_cycles, &mhz); \
        printf(", \"%s\": %.2f", label, 512.0 * (nops) * 2 * 4096 / cyc);                            \
    }
    RUN_RATIO(0, "ratio_7alu_4add_lane_ops", 22) RUN_RATIO(1, "alu_only_7_lane_ops", 14)
    RUN_RATIO(2, "mix_viaddmnmx+vimnmx3", 22) RUN_RATIO(3, "ratio_3alu_1add_lane_ops", 22)
    printf("}\n");
    return cudaGetLastError() != cudaSuccess;
}
