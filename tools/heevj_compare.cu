// heevj_compare.cu — cuSOLVER's batched Hermitian Jacobi eigensolver (cusolverDnZheevjBatched) timed on the batch shapes of
// this path, as the external yardstick for libmepb's own eigensolvers (BASELINE.json north_star: "cuSOLVER syevjBatched timed
// only as a comparison"; the op replaced is the per-k zheevd of src/matrix_math.f90:192-226).  NOT part of the product: libmepb
// never links cuSOLVER.
//
//   heevj_compare <n> <batch> [reps]      prints one JSON line: {"n":..,"batch":..,"ms":..,"matrices_per_s":..,"max_resid":..}
//
// Matrices: H = A + A^dagger with A uniform in (-0.5, 0.5)^2 plus a spread diagonal (like H(k) of a tight-binding model),
// eigenvectors requested (jobz = VECTOR, uplo = UPPER like the reference's zheevd('V','U')), tolerance 1e-14, at most 100 sweeps.
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CS(x) do { cusolverStatus_t s_ = (x); if (s_ != CUSOLVER_STATUS_SUCCESS) { std::fprintf(stderr, "%s: cusolver status %d\n", #x, (int)s_); return 1; } } while (0)

__global__ void fill_kernel(cuDoubleComplex *A, int n, long long batch, unsigned long long seed) {
    const long long total = batch * n * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / (n * n);
        const int e = (int)(i % (n * n)), r = e % n, c = e / n;
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        unsigned long long x = seed + (unsigned long long)b * 0x9E3779B97F4A7C15ull + (unsigned long long)(lo * n + hi) * 0xBF58476D1CE4E5B9ull;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        const double u = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        unsigned long long y = x * 0x9E3779B97F4A7C15ull; y ^= y >> 29;
        const double v = (double)(y >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        cuDoubleComplex z;
        if (r == c) z = make_cuDoubleComplex(2.0 * u + 3.0 * ((double)r / n - 0.5), 0.0);
        else z = make_cuDoubleComplex(u, r < c ? v : -v);
        A[i] = z;
    }
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 8;
    const long long batch = argc > 2 ? std::atoll(argv[2]) : (1 << 21);
    const int reps = argc > 3 ? std::atoi(argv[3]) : 3;
    cusolverDnHandle_t h;
    CS(cusolverDnCreate(&h));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    CS(cusolverDnSetStream(h, st));
    syevjInfo_t par;
    CS(cusolverDnCreateSyevjInfo(&par));
    CS(cusolverDnXsyevjSetTolerance(par, 1e-14));
    CS(cusolverDnXsyevjSetMaxSweeps(par, 100));
    CS(cusolverDnXsyevjSetSortEig(par, 1));
    cuDoubleComplex *A = nullptr, *A0 = nullptr, *work = nullptr;
    double *W = nullptr;
    int *info = nullptr;
    const size_t elems = (size_t)batch * n * n;
    CK(cudaMalloc(&A, sizeof(cuDoubleComplex) * elems));
    CK(cudaMalloc(&A0, sizeof(cuDoubleComplex) * elems));
    CK(cudaMalloc(&W, sizeof(double) * batch * n));
    CK(cudaMalloc(&info, sizeof(int) * batch));
    fill_kernel<<<148 * 8, 256, 0, st>>>(A0, n, batch, 20261018ull);
    CK(cudaGetLastError());
    int lwork = 0;
    CS(cusolverDnZheevjBatched_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, W, &lwork, par, (int)batch));
    CK(cudaMalloc(&work, sizeof(cuDoubleComplex) * (size_t)lwork));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {                    // first repetition is the warm-up
        CK(cudaMemcpyAsync(A, A0, sizeof(cuDoubleComplex) * elems, cudaMemcpyDeviceToDevice, st));
        CK(cudaEventRecord(e0, st));
        CS(cusolverDnZheevjBatched(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, A, n, W, work, lwork, info, par, (int)batch));
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    // residual ||A0 V - V diag(W)||_max on the first few matrices, and convergence flags
    const int ncheck = batch < 4 ? (int)batch : 4;
    std::vector<cuDoubleComplex> hA((size_t)ncheck * n * n), hV((size_t)ncheck * n * n);
    std::vector<double> hW((size_t)ncheck * n);
    std::vector<int> hinfo((size_t)batch);
    CK(cudaMemcpy(hA.data(), A0, sizeof(cuDoubleComplex) * hA.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hV.data(), A, sizeof(cuDoubleComplex) * hV.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hW.data(), W, sizeof(double) * hW.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hinfo.data(), info, sizeof(int) * batch, cudaMemcpyDeviceToHost));
    long long nfail = 0;
    for (long long b = 0; b < batch; ++b) nfail += hinfo[b] != 0;
    double resid = 0.0;
    for (int b = 0; b < ncheck; ++b)
        for (int j = 0; j < n; ++j)
            for (int r = 0; r < n; ++r) {
                double sr = 0.0, si = 0.0;
                for (int c = 0; c < n; ++c) {
                    // upper triangle is the matrix (uplo = U): element (r, c) for r <= c, conj of (c, r) otherwise
                    cuDoubleComplex a = r <= c ? hA[(size_t)b * n * n + r + (size_t)n * c] : hA[(size_t)b * n * n + c + (size_t)n * r];
                    if (r > c) a.y = -a.y;
                    const cuDoubleComplex v = hV[(size_t)b * n * n + c + (size_t)n * j];
                    sr += a.x * v.x - a.y * v.y; si += a.x * v.y + a.y * v.x;
                }
                const cuDoubleComplex v = hV[(size_t)b * n * n + r + (size_t)n * j];
                sr -= hW[(size_t)b * n + j] * v.x; si -= hW[(size_t)b * n + j] * v.y;
                resid = std::fmax(resid, std::fmax(std::fabs(sr), std::fabs(si)));
            }
    std::printf("{\"solver\": \"cusolverDnZheevjBatched\", \"n\": %d, \"batch\": %lld, \"ms\": %.4f, \"matrices_per_s\": %.6g, \"max_resid\": %.3e, "
                "\"not_converged\": %lld, \"tol\": 1e-14, \"max_sweeps\": 100}\n",
                n, batch, best, (double)batch / (best * 1e-3), resid, nfail);
    return 0;
}
