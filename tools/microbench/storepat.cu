// storepat.cu -- which CL layout lets a whole-tree kernel's stores run at the HBM write peak?
//
// A whole-tree CL kernel is store-bound for 4 states: every CTA owns a tile of patterns and, step after
// step (node after node), writes K = 16 rows of that tile.  This microbenchmark issues exactly those
// stores -- no arithmetic, no loads -- for three layouts of a node's CL buffer:
//   rows   [k][ps]               the reference's index order; a CTA's 16 rows of a step are 16 chunks, ps*8 B apart
//   tileC  [tile][k][TC]         TC = the CTA's patterns: a CTA's step is ONE contiguous 16*TC*8-byte block
//   tileW  [tile][k][64]         64 = a warp's patterns: a warp's step is one contiguous 8 KB block
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o storepat storepat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int LAYOUT, int THREADS>
__global__ void __launch_bounds__(THREADS) wr_tree(double *base, size_t ps, int nodes, double v)
{
    constexpr int K = 16;
    const size_t pat = ((size_t)blockIdx.x * THREADS + threadIdx.x) * 2;
    if (pat >= ps) return;
    size_t off, rowStride;
    if (LAYOUT == 0) { off = pat; rowStride = ps; }
    else if (LAYOUT == 1) { const size_t T = THREADS * 2; off = (pat / T) * (K * T) + pat % T; rowStride = T; }
    else { const size_t T = 64; off = (pat / T) * (K * T) + pat % T; rowStride = T; }
    for (int n = 0; n < nodes; n++) {
        double *p = base + (size_t)n * K * ps + off;
#pragma unroll
        for (int k = 0; k < K; k++) *reinterpret_cast<double2 *>(p + k * rowStride) = make_double2(v + n, v + k);
    }
}

template <class F> float timeit(F f, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

template <int LAYOUT, int THREADS> void run(const char *name, double *p, size_t ps, int nodes)
{
    const int blocks = (int)((ps / 2 + THREADS - 1) / THREADS);
    float t = timeit([&] { wr_tree<LAYOUT, THREADS><<<blocks, THREADS>>>(p, ps, nodes, 1.0); }, 5);
    printf("%-6s threads %3d  ps %8zu  nodes %3d : %7.3f ms  %6.0f GB/s\n", name, THREADS, ps, nodes, t, nodes * 16.0 * ps * 8 / t / 1e6);
}

int main()
{
    const int nodes = 96;
    for (size_t ps : {(size_t)1000000 / 64 * 64, (size_t)250000 / 64 * 64 + 64, (size_t)125000 / 64 * 64 + 64}) {
        const size_t psr = (ps + 255) / 256 * 256;
        double *p;
        if (cudaMalloc(&p, (size_t)nodes * 16 * psr * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
        run<0, 128>("rows", p, psr, nodes);
        run<1, 128>("tileC", p, psr, nodes);
        run<2, 128>("tileW", p, psr, nodes);
        run<0, 64>("rows", p, psr, nodes);
        run<1, 64>("tileC", p, psr, nodes);
        run<0, 32>("rows", p, psr, nodes);
        run<1, 32>("tileC", p, psr, nodes);
        run<0, 256>("rows", p, psr, nodes);
        run<1, 256>("tileC", p, psr, nodes);
        cudaFree(p);
    }
    return 0;
}
