// Does a DFMA stream with three distinct register operands (the band-sum inner loops: acc[q] += w[q] * g) reach the
// DFMA peak that `acc = fma(acc, a, b)` (two loop constants) reaches?  And the same contraction issued as DMMA m8n8k4?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_operands.cu -o tools/fp64_operands
#include <cstdio>
#include <cuda_runtime.h>

// 18 accumulators: pr[q] += v[q]*ai, pi[q] += v[q]*inv  (sweep_ahc inner loop without the reciprocal)
__global__ void __launch_bounds__(256) dfma3_kernel(double* out, const double* in, int iters) {
    double v[9], pr[9], pi[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { v[q] = in[q] + threadIdx.x * 1e-9; pr[q] = 0; pi[q] = 0; }
    double ai = in[9], inv = in[10];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 9; ++q) { pr[q] = fma(v[q], ai, pr[q]); pi[q] = fma(v[q], inv, pi[q]); }
        ai += 1e-9; inv -= 1e-9;          // 2 extra FP64 instructions per 18 (counted)
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 9; ++q) s += pr[q] + pi[q];
    if (s == 123.456) out[0] = s;
}
// same, weights re-read from shared memory every iteration (broadcast LDS.128 like the staged kernel)
__global__ void __launch_bounds__(256) dfma3_lds_kernel(double* out, const double* in, int iters) {
    __shared__ double2 items[64 * 5];
    for (int i = threadIdx.x; i < 64 * 5; i += 256) items[i] = make_double2(in[i % 11], in[(i + 1) % 11]);
    __syncthreads();
    double pr[9], pi[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { pr[q] = 0; pi[q] = 0; }
    double ai = in[9] + threadIdx.x * 1e-9, inv = in[10];
    for (int it = 0; it < iters; ++it) {
        const double2* itp = items + (it & 63) * 5;
        double v[10];
#pragma unroll
        for (int q = 0; q < 5; ++q) { const double2 t = itp[q]; v[2 * q] = t.x; v[2 * q + 1] = t.y; }
#pragma unroll
        for (int q = 0; q < 9; ++q) { pr[q] = fma(v[1 + q], ai, pr[q]); pi[q] = fma(v[1 + q], inv, pi[q]); }
        ai += v[0]; inv -= 1e-9;
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 9; ++q) s += pr[q] + pi[q];
    if (s == 123.456) out[0] = s;
}
// the contraction as DMMA: per step 4 items x 8 frequencies; A = weights (LDS.64 per lane), B = per-lane generated
template <int NT>
__global__ void __launch_bounds__(256) dmma_sweep_kernel(double* out, const double* in, int iters) {
    __shared__ double items[64 * 4 * 10];
    for (int i = threadIdx.x; i < 64 * 40; i += 256) items[i] = in[i % 11];
    __syncthreads();
    const int lane = threadIdx.x & 31, m = lane >> 2, k = lane & 3;
    double c1[NT][2], c2[NT][2], ai[NT], inv[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) { c1[t][0] = c1[t][1] = c2[t][0] = c2[t][1] = 0; ai[t] = in[9] + lane * 1e-9 + t; inv[t] = in[10] + t; }
    for (int it = 0; it < iters; ++it) {
        const double a = items[(it & 63) * 40 + k * 10 + 1 + m];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c1[t][0]), "+d"(c1[t][1]) : "d"(a), "d"(ai[t]));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c2[t][0]), "+d"(c2[t][1]) : "d"(a), "d"(inv[t]));
            ai[t] += 1e-9; inv[t] -= 1e-9;
        }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < NT; ++t) s += c1[t][0] + c1[t][1] + c2[t][0] + c2[t][1];
    if (s == 123.456) out[0] = s;
}

template <typename F> float best_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out, *in; cudaMalloc(&out, 1024); cudaMalloc(&in, 1024);
    double h[16]; for (int i = 0; i < 16; ++i) h[i] = 0.1 + 0.01 * i; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 20000;
    printf("{");
    for (int bps : {2, 3, 4, 8}) {
        const int blocks = p.multiProcessorCount * bps;
        float ms = best_ms([&] { dfma3_kernel<<<blocks, 256>>>(out, in, iters); });
        printf("\"dfma3_regs_%dcta\": %.2f, ", bps, 2.0 * 20 * iters * 256.0 * blocks / ms * 1e-9);
        ms = best_ms([&] { dfma3_lds_kernel<<<blocks, 256>>>(out, in, iters); });
        printf("\"dfma3_lds_%dcta\": %.2f, ", bps, 2.0 * 20 * iters * 256.0 * blocks / ms * 1e-9);
        ms = best_ms([&] { dmma_sweep_kernel<2><<<blocks, 256>>>(out, in, iters); });
        printf("\"dmma_nt2_%dcta\": %.2f, ", bps, (2.0 * 256 * 2 + 2.0 * 2 * 32) * 2 * iters * 8.0 * blocks / ms * 1e-9);
        ms = best_ms([&] { dmma_sweep_kernel<4><<<blocks, 256>>>(out, in, iters); });
        printf("\"dmma_nt4_%dcta\": %.2f, ", bps, (2.0 * 256 * 2 + 2.0 * 2 * 32) * 4 * iters * 8.0 * blocks / ms * 1e-9);
    }
    printf("\"unit\": \"TFLOP/s (FMA = 2), peak 37.1\"}\n");
    return 0;
}
