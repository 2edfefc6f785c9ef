// calib.cu — on-box calibration of the two rooflines the svd_gpu() path is judged against:
//   * FP64 tensor (DMMA.8x8x4) and FP64 FMA peak, which MEASURED_PEAKS.json does not carry;
//   * streaming READ bandwidth (the bidiagonalization passes only read the trailing matrix).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/calib bench/calib.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void dmma_loop(double *out, int iters)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_loop(double *out, int iters)
{
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void read_stream(const double2 *__restrict__ x, size_t n2, double *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double acc = 0;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 v0 = x[i], v1 = x[i + stride], v2 = x[i + 2 * stride], v3 = x[i + 3 * stride];
        acc += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
    }
    for (; i < n2; i += stride) { double2 v = x[i]; acc += v.x + v.y; }
    if (acc == 12345.678) out[0] = acc;
}
__global__ void copy_stream(const double2 *__restrict__ x, double2 *__restrict__ y, size_t n2)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) y[i] = x[i];
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000, blocks = nsm * 2, threads = warps * 32 / 2;
        dmma_loop<<<blocks, threads>>>(out, 100); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); dmma_loop<<<blocks, threads>>>(out, iters); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (double)blocks * (threads / 32) * iters * 8 * 512.0;
        printf("{\"probe\":\"dmma\",\"warps_per_sm\":%d,\"tflops\":%.2f}\n", warps, flops / ms * 1e-9);
    }
    for (int warps = 8; warps <= 32; warps *= 2) {
        int iters = 20000, blocks = nsm * 2, threads = warps * 32 / 2;
        dfma_loop<<<blocks, threads>>>(out, 100); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); dfma_loop<<<blocks, threads>>>(out, iters); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (double)blocks * threads * iters * 16 * 2.0;
        printf("{\"probe\":\"dfma\",\"warps_per_sm\":%d,\"tflops\":%.2f}\n", warps, flops / ms * 1e-9);
    }
    size_t bytes = (size_t)4 << 30;
    double2 *x, *y; CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&y, bytes));
    CK(cudaMemset(x, 0, bytes)); CK(cudaMemset(y, 0, bytes));
    for (int rep = 0; rep < 2; ++rep)
    for (int cps = 2; cps <= 16; cps *= 2) {
        read_stream<<<nsm * cps, 256>>>(x, bytes / 16, out); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); read_stream<<<nsm * cps, 256>>>(x, bytes / 16, out); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) printf("{\"probe\":\"hbm_read\",\"ctas_per_sm\":%d,\"gbs\":%.1f}\n", cps, bytes / ms * 1e-6);
    }
    copy_stream<<<nsm * 8, 256>>>(x, y, bytes / 16); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); copy_stream<<<nsm * 8, 256>>>(x, y, bytes / 16); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"probe\":\"hbm_copy\",\"gbs\":%.1f}\n", 2.0 * bytes / ms * 1e-6);
    return 0;
}
