// Do DFMA and DMMA (mma.sync f64) share one pipe on B200?  Interleave independent chains of both in one kernel and compare
// the combined rate with the separate peaks of tools/fp64_peaks.cu (37.1 TF each).  ~37 TF combined -> one pipe.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_mixed.cu -o tools/fp64_mixed
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NM>
__global__ void __launch_bounds__(256) mixed_kernel(double* out, double a, double b, int iters) {
    double acc[NF > 0 ? NF : 1];
    double c[NM > 0 ? NM : 1][2];
    double ma = threadIdx.x * 1e-3, mb = 1.0 - threadIdx.x * 1e-3;
#pragma unroll
    for (int i = 0; i < NF; ++i) acc[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < NM; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
            if (i < NM)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(ma), "d"(mb));
            if (i < NF) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NF, int NM>
void run(double* out, int blocks, const char* name) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mixed_kernel<NF, NM><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); mixed_kernel<NF, NM><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double f_fma = 2.0 * NF * iters * 256.0 * blocks, f_mma = 2.0 * 8 * 8 * 4 * (double)NM * iters * (256 / 32) * blocks;
    printf("\"%s\": {\"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"sum\": %.2f}", name, f_fma / best * 1e-9, f_mma / best * 1e-9,
           (f_fma + f_mma) / best * 1e-9);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, 1024);
    const int blocks = p.multiProcessorCount * 8;
    printf("{");
    run<8, 0>(out, blocks, "dfma8"); printf(", ");
    run<0, 8>(out, blocks, "dmma8"); printf(", ");
    run<8, 8>(out, blocks, "dfma8_dmma8"); printf(", ");
    run<8, 4>(out, blocks, "dfma8_dmma4"); printf(", ");
    run<4, 8>(out, blocks, "dfma4_dmma8");
    printf("}\n");
    return 0;
}
