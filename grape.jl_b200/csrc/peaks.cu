// peaks.cu -- measures the FP64 roofline denominators on the box the bench runs on:
// DFMA (FP64 FMA pipe) and DMMA (mma.sync.m8n8k4.f64 tensor path) peak TFLOP/s.
// MEASURED_PEAKS.json carries HBM and bf16 only; SURVEY.md 8d asks for the FP64
// peak to be measured before any FP64 fraction is quoted.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = 1.0 + x0, x2 = 2.0 + x0, x3 = 3.0 + x0, x4 = 4.0 + x0, x5 = 5.0 + x0,
           x6 = 6.0 + x0, x7 = 7.0 + x0;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0}, c2[2] = {0.0, 0.0}, c3[2] = {0.0, 0.0};
    double av = a + threadIdx.x * 1e-9, bv = b;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[0]), "+d"(c0[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c1[0]), "+d"(c1[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c2[0]), "+d"(c2[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c3[0]), "+d"(c3[1]) : "d"(av), "d"(bv));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

static double run(int which, int device) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 20000;
    double* out;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
        else dmma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops;
        if (which == 0) flops = 2.0 * 8 * (double)iters * blocks * threads;
        else flops = 2.0 * 8 * 8 * 4 * 4 * (double)iters * blocks * (threads / 32);   // 4 mma of 8x8x4 per warp-iter
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

extern "C" double gb_peak_dfma_tflops(int device) { return run(0, device); }
extern "C" double gb_peak_dmma_tflops(int device) { return run(1, device); }
