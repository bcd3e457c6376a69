// fp64_peak.cu -- measured FP64 (non-tensor) throughput of the device: the roofline denominator of
// k_superpose ("measure with a pure-FMA FP64 microbenchmark on the box").
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_peak tools/gpu/fp64_peak.cu && /tmp/fp64_peak
// Prints one JSON line: DFMA rate (2 flop each) and the DADD+DMUL rate the engine's -fmad=false code can reach.
#include <cstdio>
#include <cuda_runtime.h>

template <bool FMA> __global__ void __launch_bounds__(256) k_flops(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (FMA) x[i] = __fma_rn(x[i], a, b);
            else x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 123.456) out[0] = s;   // never true: keeps the loop alive
}

template <bool FMA> static double run(int sms, int iters)
{
    double *out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8;
    k_flops<FMA><<<grid, 256>>>(out, iters / 10, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_flops<FMA><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 8.0 * (double)iters * 256.0 * grid;
        best = flop / (ms * 1e-3) / 1e12 > best ? flop / (ms * 1e-3) / 1e12 : best;
    }
    cudaFree(out);
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double fma = run<true>(p.multiProcessorCount, 200000), addmul = run<false>(p.multiProcessorCount, 200000);
    printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"fp64_mul_add_tflops\": %.2f, "
           "\"how\": \"8 independent chains per thread, 256 threads x 8 CTAs per SM, best of 5, CUDA events\"}\n",
           p.name, p.multiProcessorCount, fma, addmul);
    return 0;
}
