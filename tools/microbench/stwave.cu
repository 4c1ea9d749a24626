// stwave.cu -- how many L1 data-pipe wavefronts does a warp-wide global store cost, by shape?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stwave stwave.cu
// Run under: ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,smsp__inst_executed_op_global_st.sum,gpu__time_duration.sum ./stwave
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: st.v2.f64, a warp writes 512 contiguous bytes        MODE 1: st.f64, 256 contiguous bytes
// MODE 2: st.v2.f64, 4 rows x 128 B (lane = 4g+q: row q)       MODE 3: st.f64, 4 rows x 64 B
// MODE 4: st.v2.f64, 8 rows x 64 B                             MODE 5: st.v4.f64 (256-bit), 4 rows x 256 B
template <int MODE>
__global__ void k(double *out, size_t rowStride, int iters)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (int it = 0; it < iters; it++) {
        double *base = out + (size_t)it * 64 * rowStride + warp * 64;
        const double v = (double)(it + lane);
        if (MODE == 0) *reinterpret_cast<double2 *>(base + lane * 2) = make_double2(v, v);
        if (MODE == 1) base[lane] = v;
        if (MODE == 2) *reinterpret_cast<double2 *>(base + q * rowStride + g * 2) = make_double2(v, v);
        if (MODE == 3) base[q * rowStride + g] = v;
        if (MODE == 4) *reinterpret_cast<double2 *>(base + (lane >> 2) * rowStride + q * 2) = make_double2(v, v);
        if (MODE == 5) {
            double *p = base + q * rowStride + g * 4;
            asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(v) : "memory");
        }
    }
}

int main()
{
    const size_t rowStride = 1 << 18;          // doubles: rows 2 MB apart
    double *d;
    if (cudaMalloc(&d, ((size_t)32 * 64 + 16) * rowStride * 8 + (64ull << 20)) != cudaSuccess) { printf("malloc failed\n"); return 1; }
    const int blocks = 148 * 4, iters = 32;
    k<0><<<blocks, 256>>>(d, rowStride, iters);
    k<1><<<blocks, 256>>>(d, rowStride, iters);
    k<2><<<blocks, 256>>>(d, rowStride, iters);
    k<3><<<blocks, 256>>>(d, rowStride, iters);
    k<4><<<blocks, 256>>>(d, rowStride, iters);
    k<5><<<blocks, 256>>>(d, rowStride, iters);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
