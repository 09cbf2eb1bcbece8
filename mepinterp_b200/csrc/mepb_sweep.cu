// mepb_sweep.cu — frequency-sweep kernels: consume the "pair items" written by the interpolation kernels and
// accumulate velo_ahc_tens / kubo_ohc_tens (kubo.f90:86-198) and photo_2nd_cond (photo.f90:16-86) for every hw.
//
// One thread owns one frequency (and one of G k-groups); its accumulators live in registers for the whole launch:
//   out[comp, w] += sum_items weight[item, comp] * g[item](w)
// i.e. a (components x items) . (items x n_hw) contraction whose right-hand factor is generated on the fly
// (one complex reciprocal per (item, w)).  FP64 FMA pipe bound; items are read once from HBM/L2 (broadcast loads).
#include "mepb_internal.cuh"

namespace mepb {

namespace {

// ------------------------------------------------------------------------------------------------ AHC / OHC sweep
// HERM: Hermitian items (10 doubles) -> accumulators X2 (3 complex), Y1 (6 complex)            = 18 doubles
// else: general items (20 doubles)   -> accumulators X1, X2 (3 complex each), Y1, Y2 (6 each) = 36 doubles
template <bool HERM>
__global__ void __launch_bounds__(256, 2)
sweep_ahc_kernel(const double2 *__restrict__ items, const int *__restrict__ cnt, const int max_pairs, const long long nk,
                 const int n_hw, const double *__restrict__ hw, const double eta, double *__restrict__ block_out,
                 const int G, const int NS) {
    constexpr int NACC = HERM ? kSweepAccH : kSweepAcc;
    constexpr int ISTR = (HERM ? kAhcItemH : kAhcItemG) / 2;     // item stride in double2
    const long long per = (nk + gridDim.x - 1) / gridDim.x;
    const long long kb0 = (long long)blockIdx.x * per;
    const long long kb1 = (kb0 + per < nk) ? kb0 + per : nk;
    for (int s = threadIdx.x; s < NS; s += blockDim.x) {
        const int w = s % n_hw, g = s / n_hw;
        const double om = hw[w];
        const double om2 = om * om, b = -2.0 * om * eta, b2 = b * b;
        double acc[NACC];
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[q] = 0.0;
        for (long long k = kb0 + g; k < kb1; k += G) {
            const int c = cnt[k];
            const double2 *it = items + (size_t)k * max_pairs * ISTR;
#pragma unroll 2
            for (int j = 0; j < c; ++j, it += ISTR) {
                double v[2 * ISTR];
#pragma unroll
                for (int q = 0; q < ISTR; ++q) { const double2 t = __ldg(it + q); v[2 * q] = t.x; v[2 * q + 1] = t.y; }
                // g = 1/(dE^2 - (w + i eta)^2) = 1/(a + i b),  a = dE^2 + eta^2 - w^2,  b = -2 w eta
                const double a = v[0] - om2;
                const double inv = fast_rcp(fma(a, a, b2));
                const double gr = a * inv, gi = -b * inv;
                const double zr = om * gr - eta * gi, zi = fma(om, gi, eta * gr);      // z g
                if (HERM) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { acc[2 * q] = fma(v[1 + q], gr, acc[2 * q]); acc[2 * q + 1] = fma(v[1 + q], gi, acc[2 * q + 1]); }
#pragma unroll
                    for (int q = 0; q < 6; ++q) { acc[6 + 2 * q] = fma(v[4 + q], zr, acc[6 + 2 * q]); acc[7 + 2 * q] = fma(v[4 + q], zi, acc[7 + 2 * q]); }
                } else {
#pragma unroll
                    for (int q = 0; q < 6; ++q) { acc[2 * q] = fma(v[1 + q], gr, acc[2 * q]); acc[2 * q + 1] = fma(v[1 + q], gi, acc[2 * q + 1]); }
#pragma unroll
                    for (int q = 0; q < 12; ++q) { acc[12 + 2 * q] = fma(v[7 + q], zr, acc[12 + 2 * q]); acc[13 + 2 * q] = fma(v[7 + q], zi, acc[13 + 2 * q]); }
                }
            }
        }
        double *o = block_out + ((size_t)blockIdx.x * NS + s) * NACC;
#pragma unroll
        for (int q = 0; q < NACC; ++q) o[q] = acc[q];
    }
}

// ------------------------------------------------------------------------------------------------ photo sweep
__global__ void __launch_bounds__(128, 2)
sweep_photo_kernel(const double2 *__restrict__ items, const int *__restrict__ cnt, const int max_items, const long long nk,
                   const int n_hw, const double *__restrict__ hw, const double eta, double *__restrict__ block_out,
                   const int G, const int NS) {
    const long long per = (nk + gridDim.x - 1) / gridDim.x;
    const long long kb0 = (long long)blockIdx.x * per;
    const long long kb1 = (kb0 + per < nk) ? kb0 + per : nk;
    for (int s = threadIdx.x; s < NS; s += blockDim.x) {
        const int w = s % n_hw, g = s / n_hw;
        const double om = hw[w];
        const double om2 = om * om;
        const double iw2 = 1.0 / om2;                 // hw = 0 -> inf/NaN like photo.f90:70 (Q13)
        const double e2 = 2.0 * eta;
        double2 acc[27];
#pragma unroll
        for (int q = 0; q < 27; ++q) acc[q] = make_double2(0.0, 0.0);
        for (long long k = kb0 + g; k < kb1; k += G) {
            const int c = cnt[k];
            const double2 *it = items + (size_t)k * max_items * (kPhotoItem / 2);
            for (int j = 0; j < c; ++j, it += kPhotoItem / 2) {
                const double xr = __ldg(it + 27).x;
                // h = (1/w^2) * 2x/(x^2 - w^2),  x = xr - 2 i eta
                const double a = xr * xr - e2 * e2 - om2, b = -2.0 * xr * e2;
                const double inv = fast_rcp(fma(a, a, b * b)) * iw2;
                const double hr = 2.0 * (xr * a - e2 * b) * inv, hi = -2.0 * (xr * b + e2 * a) * inv;
#pragma unroll
                for (int q = 0; q < 27; ++q) {
                    const double2 t = __ldg(it + q);
                    acc[q].x = fma(t.x, hr, acc[q].x); acc[q].x = fma(-t.y, hi, acc[q].x);
                    acc[q].y = fma(t.x, hi, acc[q].y); acc[q].y = fma(t.y, hr, acc[q].y);
                }
            }
        }
        double2 *o = reinterpret_cast<double2 *>(block_out + ((size_t)blockIdx.x * NS + s) * kPhotoAcc);
#pragma unroll
        for (int q = 0; q < 27; ++q) o[q] = acc[q];
    }
}

// ------------------------------------------------------------------------------------------------ deterministic column sums
// out_accum[(c/group)*out_stride + out_off + c%group] += sum_r in[r*cols + c]; one block per column, fixed tree order.
__global__ void __launch_bounds__(256) reduce_rows_kernel(const double *__restrict__ in, const int rows, const int cols,
                                                          double *__restrict__ out_accum, const int group,
                                                          const int out_stride, const int out_off) {
    __shared__ double sm[256];
    for (int c = blockIdx.x; c < cols; c += gridDim.x) {
        double s = 0.0;
        for (int r = threadIdx.x; r < rows; r += 256) s += in[(size_t)r * cols + c];
        sm[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out_accum[(size_t)(c / group) * out_stride + out_off + c % group] += sm[0];
        __syncthreads();
    }
}

void sweep_shape(int n_hw, int T, int *G, int *NS, int *threads) {
    int g = T / n_hw; if (g < 1) g = 1;
    *G = g; *NS = g * n_hw;
    int t = (*NS + 31) / 32 * 32; if (t > T) t = T;
    *threads = t;
}

}  // namespace

int sweep_ahc_grid(int sm_count, long long nk, int n_hw, int *slots_out) {
    int G, NS, T; sweep_shape(n_hw, 256, &G, &NS, &T);
    if (slots_out) *slots_out = NS;
    long long g = (long long)sm_count * 2;
    long long need = (nk + G - 1) / G;      // at least G k-points per block to give every group work
    if (g > need) g = need;
    if (g < 1) g = 1;
    return (int)g;
}
cudaError_t launch_sweep_ahc(bool herm, const double *items, const int *cnt, int max_pairs, long long nk, int n_hw,
                             const double *hw, double eta, double *block_out, int grid, cudaStream_t st) {
    int G, NS, T; sweep_shape(n_hw, 256, &G, &NS, &T);
    if (herm)
        sweep_ahc_kernel<true><<<grid, T, 0, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_pairs, nk, n_hw, hw, eta, block_out, G, NS);
    else
        sweep_ahc_kernel<false><<<grid, T, 0, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_pairs, nk, n_hw, hw, eta, block_out, G, NS);
    return cudaGetLastError();
}
int sweep_photo_grid(int sm_count, long long nk, int n_hw, int *slots_out) {
    int G, NS, T; sweep_shape(n_hw, 128, &G, &NS, &T);
    if (slots_out) *slots_out = NS;
    long long g = (long long)sm_count * 2;
    long long need = (nk + G - 1) / G;
    if (g > need) g = need;
    if (g < 1) g = 1;
    return (int)g;
}
cudaError_t launch_sweep_photo(const double *items, const int *cnt, int max_items, long long nk, int n_hw, const double *hw,
                               double eta, double *block_out, int grid, cudaStream_t st) {
    int G, NS, T; sweep_shape(n_hw, 128, &G, &NS, &T);
    sweep_photo_kernel<<<grid, T, 0, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_items, nk, n_hw, hw, eta, block_out, G, NS);
    return cudaGetLastError();
}
cudaError_t launch_reduce_rows(const double *in, int rows, int cols, double *out_accum, int group, int out_stride, int out_off,
                               cudaStream_t st) {
    int grid = cols < 4096 ? cols : 4096;
    reduce_rows_kernel<<<grid, 256, 0, st>>>(in, rows, cols, out_accum, group, out_stride, out_off);
    return cudaGetLastError();
}

}  // namespace mepb
