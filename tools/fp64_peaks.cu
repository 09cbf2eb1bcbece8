// FP64 roofline denominators for B200 (sm_100a): DFMA pipe, DMMA (mma.sync f64) shapes,
// cuBLAS DGEMM / ZGEMM.  MEASURED_PEAKS.json (driver-written) has no FP64 entry, so the
// Fourier-sum / band-sum kernels of this repo quote their roofline against what this tool prints.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/fp64_peaks.cu -lcublas -o tools/fp64_peaks
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cuComplex.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b, int iters) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs
template <int ILP>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters) {
    double c[ILP][2];
    double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs
template <int ILP>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters) {
    double c[ILP][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1.0 - threadIdx.x * 1e-3 * i;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// m16n8k4: A 2 regs, B 1 reg, C 4 regs
template <int ILP>
__global__ void __launch_bounds__(256) dmma1684_kernel(double* out, int iters) {
    double c[ILP][4];
    double a0 = threadIdx.x * 1e-3, a1 = 0.5 + a0, b = 1.0 - threadIdx.x * 1e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) sincospi_kernel(double* out, int iters) {
    double x = threadIdx.x * 1.234e-3 + blockIdx.x * 1e-5, s = 0, c = 0, acc = 0;
    for (int it = 0; it < iters; ++it) { sincospi(x, &s, &c); acc += s * c; x += 0.37; }
    if (acc == 123.456) out[0] = acc;
}
__global__ void __launch_bounds__(256) ddiv_kernel(double* out, int iters) {
    double x = 1.0 + threadIdx.x * 1.234e-3, acc = 0;
    for (int it = 0; it < iters; ++it) { acc += 1.0 / x; x += 0.37; }
    if (acc == 123.456) out[0] = acc;
}
__global__ void __launch_bounds__(256) drcp_kernel(double* out, int iters) {
    // rcp.approx.ftz.f64 + 2 Newton steps (what the band-sum kernels use)
    double x = 1.0 + threadIdx.x * 1.234e-3, acc = 0;
    for (int it = 0; it < iters; ++it) {
        double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r);
        acc += r; x += 0.37;
    }
    if (acc == 123.456) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, clk);
    double* out; CK(cudaMalloc(&out, 1024));
    const int iters = 20000;
    const int blocks = sms * 8;
    {
        float ms = time_ms([&] { dfma_kernel<8><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters); });
        double fl = 2.0 * 8 * iters * 256.0 * blocks;
        printf(", \"dfma_tflops\": %.3f", fl / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { dmma884_kernel<8><<<blocks, 256>>>(out, iters); });
        double fl = 2.0 * 8 * 8 * 4 * 8.0 * iters * (256 / 32) * blocks;
        printf(", \"dmma_m8n8k4_tflops\": %.3f", fl / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { dmma1684_kernel<8><<<blocks, 256>>>(out, iters); });
        double fl = 2.0 * 16 * 8 * 4 * 8.0 * iters * (256 / 32) * blocks;
        printf(", \"dmma_m16n8k4_tflops\": %.3f", fl / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { dmma16816_kernel<4><<<blocks, 256>>>(out, iters / 4); });
        double fl = 2.0 * 16 * 8 * 16 * 4.0 * (iters / 4) * (256 / 32) * blocks;
        printf(", \"dmma_m16n8k16_tflops\": %.3f", fl / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { sincospi_kernel<<<blocks, 256>>>(out, 2000); });
        printf(", \"sincospi_G_per_s\": %.3f", 2000.0 * 256 * blocks / ms * 1e-6);
    }
    {
        float ms = time_ms([&] { ddiv_kernel<<<blocks, 256>>>(out, 2000); });
        printf(", \"ddiv_G_per_s\": %.3f", 2000.0 * 256 * blocks / ms * 1e-6);
    }
    {
        float ms = time_ms([&] { drcp_kernel<<<blocks, 256>>>(out, 2000); });
        printf(", \"drcp_nr2_G_per_s\": %.3f", 2000.0 * 256 * blocks / ms * 1e-6);
    }
    // cuBLAS
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {4096, 8192}) {
        double *A, *B, *C; size_t bytes = (size_t)n * n * sizeof(double);
        CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
        double al = 1.0, be = 0.0;
        float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 3);
        printf(", \"cublas_dgemm_%d_tflops\": %.3f", n, 2.0 * n * n * (double)n / ms * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    {
        int n = 4096;
        cuDoubleComplex *A, *B, *C; size_t bytes = (size_t)n * n * sizeof(cuDoubleComplex);
        CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
        cuDoubleComplex al = make_cuDoubleComplex(1, 0), be = make_cuDoubleComplex(0, 0);
        float ms = time_ms([&] { cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 3);
        printf(", \"cublas_zgemm_%d_tflops\": %.3f", n, 8.0 * n * n * (double)n / ms * 1e-9);
        // sustained: 3 s loop
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        int reps = (int)(3000.0f / ms) + 1;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float tot; cudaEventElapsedTime(&tot, e0, e1);
        printf(", \"cublas_zgemm_%d_sustained_tflops\": %.3f", n, 8.0 * n * n * (double)n * reps / tot * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    {   // sustained DFMA 3 s
        float ms1 = time_ms([&] { dfma_kernel<8><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters); }, 1);
        int reps = (int)(3000.0f / ms1) + 1;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) dfma_kernel<8><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float tot; cudaEventElapsedTime(&tot, e0, e1);
        printf(", \"dfma_sustained_tflops\": %.3f", 2.0 * 8 * iters * 256.0 * blocks * reps / tot * 1e-9);
    }
    printf("}\n");
    return 0;
}
