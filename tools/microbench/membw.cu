// membw.cu -- HBM read / write / copy bandwidth and FP64 FMA throughput on the
// device, to put denominators under the CL-kernel roofline.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membw membw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void wr(double2 *p, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) p[i] = make_double2(v, v);
}
__global__ void rd(const double2 *p, size_t n, double *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    double a = 0;
    for (; i < n; i += s) { double2 v = p[i]; a += v.x + v.y; }
    if (a == 123.456) *out = a;
}
__global__ void cp(const double2 *p, double2 *q, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) q[i] = p[i];
}
// 16 rows written per thread at a large row stride, like a CL store
__global__ void wr_rows(double *p, size_t ps, double v)
{
    size_t pat = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (pat >= ps) return;
    for (int k = 0; k < 16; k++) *reinterpret_cast<double2 *>(p + k * ps + pat) = make_double2(v, v + k);
}
__global__ void fma64(double *out, int iters)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma64(double *out, int iters)
{
    double c0[8], c1[8];
    for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x + i; c1[i] = i; }
    const double a = 1.0000001, b = 0.9999999;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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

int main()
{
    const size_t bytes = (size_t)4 << 30;
    const size_t n = bytes / sizeof(double2);
    double2 *p, *q; double *o;
    cudaMalloc(&p, bytes); cudaMalloc(&q, bytes); cudaMalloc(&o, 1 << 24);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    for (int mult : {4, 8, 16, 32}) {
        int grid = sm * mult;
        float tw = timeit([&] { wr<<<grid, 256>>>(p, n, 1.0); }, 5);
        float tr = timeit([&] { rd<<<grid, 256>>>(p, n, o); }, 5);
        float tc = timeit([&] { cp<<<grid, 256>>>(p, q, n); }, 5);
        printf("grid=%d*SM  write %.0f GB/s  read %.0f GB/s  copy(r+w) %.0f GB/s\n", mult, bytes / tw / 1e6, bytes / tr / 1e6, 2.0 * bytes / tc / 1e6);
    }
    {
        const size_t ps = 1000000;   // 16 rows x 1M patterns = 128 MB per "node"
        const int nodes = 24;
        float t = timeit([&] { for (int k = 0; k < nodes; k++) wr_rows<<<(ps / 2 + 255) / 256, 256>>>((double *)p + (size_t)k * 16 * ps, ps, 1.0); }, 5);
        printf("CL-shaped stores (16 rows x 1M, %d launches): %.0f GB/s\n", nodes, nodes * 16.0 * ps * 8 / t / 1e6);
    }
    float tm = timeit([&] { cudaMemsetAsync(p, 0, bytes); }, 5);
    printf("cudaMemset %.0f GB/s\n", bytes / tm / 1e6);
    {
        const int iters = 4096, blocks = sm * 8, threads = 256;
        float t = timeit([&] { fma64<<<blocks, threads>>>(o, iters); }, 5);
        double flops = 2.0 * 8 * iters * (double)blocks * threads;
        printf("FP64 FMA: %.1f TFLOP/s (%.1f FMA lanes/clk/SM at 1.965 GHz)\n", flops / t / 1e9, flops / 2 / (t * 1e-3) / sm / 1.965e9);
    }
    {
        const int iters = 4096, blocks = sm * 8, threads = 256;
        float t = timeit([&] { dmma64<<<blocks, threads>>>(o, iters); }, 5);
        double flops = 2.0 * 256 * 8 * iters * (double)blocks * threads / 32;
        printf("FP64 mma.sync m8n8k4: %.1f TFLOP/s\n", flops / t / 1e9);
    }
    return 0;
}
