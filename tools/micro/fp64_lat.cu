// Dependent-chain latency and throughput of FP64 ops on the device (development helper).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_dfma(double* out, double a, double b, int n, long long* cyc)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_dadd(double* out, double a, int n, long long* cyc)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = x + a;
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_ffma(float* out, float a, float b, int n, long long* cyc)
{
    float x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = fmaf(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void tput_dfma(double* out, double a, double b, int n)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < n; ++i)
    {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main()
{
    double* d; float* f; long long* c; long long h;
    cudaMalloc(&d, 1 << 24); cudaMalloc(&f, 1 << 24); cudaMalloc(&c, 8);
    const int n = 100000;
    for (int rep = 0; rep < 2; ++rep)
    {
        chain_dfma<<<1, 32>>>(d, 1.0000001, 1e-9, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("DFMA dependent chain, 1 warp : %.2f cycles/op\n", (double)h / n);
        chain_dadd<<<1, 32>>>(d, 1e-9, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("DADD dependent chain, 1 warp : %.2f cycles/op\n", (double)h / n);
        chain_ffma<<<1, 32>>>(f, 1.0000001f, 1e-9f, n, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("FFMA dependent chain, 1 warp : %.2f cycles/op\n", (double)h / n);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks : {148, 148 * 8})
    {
        tput_dfma<<<blocks, 256>>>(d, 1.0000001, 1e-9, 20000);
        cudaEventRecord(e0); tput_dfma<<<blocks, 256>>>(d, 1.0000001, 1e-9, 20000); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * 20000.0 * blocks * 256;
        printf("DFMA throughput, %d blocks x 256: %.2f TFLOP/s (%.3f ms)\n", blocks, flops / ms / 1e9, ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
