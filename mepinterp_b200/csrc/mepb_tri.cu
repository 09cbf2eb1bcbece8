// mepb_tri.cu — Hermitian eigensolver of the general path (num_wann 17..64): Householder tridiagonalisation + implicit QL.
//
// Replaces the reference's zheevd call (matrix_math.f90:192-226) for the CTA-per-k path.  A cyclic Jacobi solver needs
// ~9 sweeps x 7.7 MFLOP at num_wann = 64 and three block barriers per rotation round; the classical dense route needs
// ~25 n^3 = 6.6 MFLOP, and its only serial part (the QL recurrence on the tridiagonal) is O(n^2) scalar work that is
// peeled off into a thread-per-k kernel where millions of k-points hide its latency:
//
//   tridiag_kernel<NE>  one CTA (8 NE threads) per k.  The trailing matrix lives in REGISTERS: thread (i, part) owns
//       A(i, part + 8 jj), jj < NE/8 — every register index is a compile-time constant (the step loop is unrolled over
//       the 8-column blocks).  Per column: zlarfg reflector, p = tau A v (8 complex FMAs + 3 shuffle stages),
//       w = p - tau/2 (p^H v) v, rank-2 update A -= v w^H + w v^H.  Three block barriers per column.
//       Same arithmetic as LAPACK zhetd2('L'): T = Q^H A Q real symmetric tridiagonal, Q = H_0 H_1 ... H_{n-2}.
//   tql_kernel          one THREAD per k: implicit-shift QL (EISPACK tql2 recurrence) on (d, e) held in shared memory
//       ([index][thread], conflict free).  Writes the eigenvalues (sorted, plus the sort ranks) and a LOG of the plane
//       rotations (c, s) with one header word per chase instead of touching any eigenvector.
//   zrot_kernel<NE>     replays the log on Z = 1.  One thread per ROW of Z, the whole row in registers (plane indices are
//       compile-time: the chase is unrolled over all planes and skipped in warp-uniform blocks of 8), so a rotation is
//       one uniform 16-byte load and four FP64 instructions, no shared memory, no barriers.
//   The reflectors are applied to Z (U = Q Z) at the head of eig_rot_large_kernel (mepb_large.cu).
#include "mepb_internal.cuh"

namespace mepb {

namespace {

constexpr int kTqlThreads = 32;

__device__ __forceinline__ double2 shfl_xor2(double2 v, int o) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}

// ================================================================================================ tridiagonalisation
// PARTS threads per matrix row (thread (i, part) owns A(i, part + PARTS jj)); NE = 64: PARTS = 4 -> 256 threads and 16
// columns per thread, ~100 registers, TWO matrices resident per SM (the 512-thread / 124-register shape of the first
// version allowed one, and the kernel is bound by its three barriers per column, not by arithmetic).  The reflector is never
// materialised per thread: with v = scale x (x = the published column, v(s+1) = 1)
//   A v = scale sum_{j>s+1} A(:,j) x_j + A(:,s+1),      w_i conj(v_j) = (w_i conj(scale)) conj(x_j)
// so the column entries come straight from shared memory (one complex multiply per thread instead of one per element).
template <int NE, int PARTS>
__global__ void __launch_bounds__(NE * PARTS, 2)
tridiag_kernel(const int nw, const long long nk, const double2 *__restrict__ Hk_in, const TriBuf tb) {
    constexpr int JJ = NE / PARTS, NT = NE * PARTS, NWARP = NT / 32;
    __shared__ double2 xs[2][NE];
    __shared__ double2 ws[NE];
    __shared__ double2 red[NWARP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = tid / PARTS, part = tid % PARTS;
    const int nw2 = nw * nw;

    for (long long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        const double2 *Hk = Hk_in + (size_t)ik * nw2;
        double2 *Vr = tb.Vr + (size_t)ik * NE * NE;
        double2 *tau_o = tb.tau + (size_t)ik * NE;
        double *d_o = tb.d + (size_t)ik * NE, *e_o = tb.e + (size_t)ik * NE;
        // zheevd 'U' semantics: only the upper triangle of H(k) is referenced, Im of the diagonal ignored
        double2 a[JJ];
#pragma unroll
        for (int jj = 0; jj < JJ; ++jj) {
            const int j = part + PARTS * jj;
            double2 v = make_double2(0.0, 0.0);
            if (i < nw && j < nw) {
                if (i < j) v = Hk[i + nw * j];
                else if (i > j) v = cconj(Hk[j + nw * i]);
                else v = make_double2(Hk[i + nw * i].x, 0.0);
            }
            a[jj] = v;
        }
        if (tid < NE) { tau_o[tid] = make_double2(0.0, 0.0); e_o[tid] = 0.0; }
        __syncthreads();     // previous matrix is done with xs / ws / red

#pragma unroll
        for (int sb = 0; sb < JJ; ++sb) {
            for (int sl = 0; sl < PARTS; ++sl) {
                const int s = PARTS * sb + sl;
                if (s > nw - 1) break;
                const int buf = s & 1;
                if (part == sl) {
                    xs[buf][i] = (i > s) ? a[sb] : make_double2(0.0, 0.0);
                    if (i == s) d_o[s] = a[sb].x;
                }
                if (s == nw - 1) break;
                __syncthreads();                                                            // (1) x visible
                // Warps whose rows are all finished (i <= s) have nothing to add: reflector entry 0, p = 0, no update.  They only keep
                // the barriers (and a zero partial of p^H v); the issue slots they leave go to the co-resident CTA.  Half of the
                // warps are in that state on average — the trailing matrix shrinks, the CTA does not.
                const bool idle = (warp * (32 / PARTS) + (32 / PARTS) - 1) <= s;
                double2 tau = make_double2(0.0, 0.0), scale = tau, vi = tau, p = tau;
                if (!idle) {
                    const double2 alpha = xs[buf][s + 1];
                    double pn = 0.0;
#pragma unroll
                    for (int r = lane; r < NE; r += 32)
                        if (r > s + 1) { const double2 x = xs[buf][r]; pn = fma(x.x, x.x, fma(x.y, x.y, pn)); }
                    const double xnorm2 = warp_sum(pn);
                    double beta;
                    if (xnorm2 == 0.0 && alpha.y == 0.0) {                                      // zlarfg: H = 1
                        tau = make_double2(0.0, 0.0); scale = make_double2(0.0, 0.0); beta = alpha.x;
                    } else {
                        beta = -copysign(sqrt(fma(alpha.x, alpha.x, fma(alpha.y, alpha.y, xnorm2))), alpha.x);
                        const double rb = 1.0 / beta;
                        tau = make_double2((beta - alpha.x) * rb, -alpha.y * rb);
                        scale = crcp(make_double2(alpha.x - beta, alpha.y));       // (MUFU-seeded forms of this chain: no gain, r03y)
                    }
                    // own row's reflector entry
                    vi = (i <= s) ? make_double2(0.0, 0.0) : (i == s + 1) ? make_double2(1.0, 0.0) : cmul(xs[buf][i], scale);
                    // p = tau A v  (rows <= s are finished: p = 0 there)
                    double2 p1 = make_double2(0.0, 0.0), p2 = make_double2(0.0, 0.0);
#pragma unroll
                    for (int jj = 0; jj < JJ; ++jj) {
                        const int j = part + PARTS * jj;
                        if (j > s + 1) cfma(p1, a[jj], xs[buf][j]);
                        else if (j == s + 1) p2 = a[jj];
                    }
                    p = cadd(cmul(scale, p1), p2);
#pragma unroll
                    for (int o = 1; o < PARTS; o <<= 1) p = cadd(p, shfl_xor2(p, o));
                    p = cmul(tau, p);
                    if (i <= s) p = make_double2(0.0, 0.0);
                    // p^H v
                    double2 dt = (part == 0) ? cmulc(vi, p) : make_double2(0.0, 0.0);
#pragma unroll
                    for (int o = PARTS; o < 32; o <<= 1) dt = cadd(dt, shfl_xor2(dt, o));
                    if (lane == 0) red[warp] = dt;
                    if (part == 0) Vr[(size_t)s * NE + i] = vi;      // (rows of idle warps keep the 0 the buffer was initialised with)
                    if (i == s + 1 && part == 0) { tau_o[s] = tau; e_o[s] = beta; }
                } else if (lane == 0) red[warp] = make_double2(0.0, 0.0);
                __syncthreads();                                                            // (2)
                double2 wi = make_double2(0.0, 0.0);
                if (!idle) {
                    double2 dot = make_double2(0.0, 0.0);
#pragma unroll
                    for (int w = 0; w < NWARP; ++w) dot = cadd(dot, red[w]);
                    const double2 alc = cscale(-0.5, cmul(tau, dot));
                    wi = p;
                    cfma(wi, alc, vi);
                    if (part == 0) ws[i] = wi;
                }
                __syncthreads();                                                            // (3)
                if (!idle) {
                    const double2 wis = cmulc(wi, scale);                                   // w_i conj(scale)
#pragma unroll
                    for (int jj = 0; jj < JJ; ++jj) {
                        const int j = part + PARTS * jj;
                        if (j > s) {
                            const double2 u1 = cmulc(vi, ws[j]);
                            const double2 u2 = (j == s + 1) ? wi : cmulc(wis, xs[buf][j]);
                            a[jj].x -= u1.x + u2.x; a[jj].y -= u1.y + u2.y;
                        }
                    }
                }
            }
        }
    }
}

// ================================================================================================ implicit QL, thread per k
__global__ void __launch_bounds__(kTqlThreads)
tql_kernel(const int n, const long long nk, const TriBuf tb, int *__restrict__ err_flag) {
    extern __shared__ double sm[];
    const int ne = tb.ne, t = threadIdx.x;
    const long long ik = (long long)blockIdx.x * kTqlThreads + t;
    if (ik >= nk) return;
    double *dd = sm + t, *ee = sm + (size_t)ne * kTqlThreads + t;
#define D(i) dd[(i) * kTqlThreads]
#define E(i) ee[(i) * kTqlThreads]
    {
        const double *dg = tb.d + (size_t)ik * ne, *eg = tb.e + (size_t)ik * ne;
        for (int i = 0; i < n; ++i) { D(i) = dg[i]; E(i) = (i < n - 1) ? eg[i] : 0.0; }
    }
    unsigned int *hdr = tb.hdr + (size_t)ik * tb.hdr_cap;
    double2 *lg = tb.log + (size_t)ik * tb.log_cap;
    const double eps = 1.1102230246251565e-16;
    int nh = 0, pos = 0;
    bool failed = false;
    for (int l = 0; l < n && !failed; ++l) {
        int iter = 0;
        while (true) {
            int m = l;
            for (; m < n - 1; ++m) {
                const double tst = fabs(D(m)) + fabs(D(m + 1));
                if (fabs(E(m)) <= eps * tst) break;
            }
            if (m == l) break;
            if (++iter > 60 || nh >= tb.hdr_cap || pos + (m - l) > tb.log_cap) { failed = true; break; }
            double g = (D(l + 1) - D(l)) / (2.0 * E(l));
            double r = sqrt(fma(g, g, 1.0));
            g = D(m) - D(l) + E(l) / (g + copysign(r, g));
            double s = 1.0, c = 1.0, p = 0.0;
            int i = m - 1;
            for (; i >= l; --i) {
                const double ei = E(i);
                const double f = s * ei, b = c * ei;
                // r = sqrt(f^2 + g^2) and 1/r from ONE MUFU-seeded reciprocal square root (the IEEE sqrt + division they replace were
                // two thirds of the dependent chain of a plane, and this kernel is nothing but that chain: ncu r03v, 55 % `wait`)
                const double h2 = fma(f, f, g * g);
                double rr;
                if (h2 > 1e-280) { rr = fast_rsqrt3(h2); r = h2 * rr; }
                else { r = sqrt(h2); rr = (r == 0.0) ? 0.0 : 1.0 / r; }
                E(i + 1) = r;
                if (r == 0.0) { D(i + 1) -= p; E(m) = 0.0; break; }
                s = f * rr; c = g * rr;
                g = D(i + 1) - p;
                r = (D(i) - g) * s + 2.0 * c * b;
                p = s * r;
                D(i + 1) = g + p;
                g = c * r - b;
                lg[pos++] = make_double2(c, s);
            }
            hdr[nh++] = (unsigned int)(i + 1) | ((unsigned int)m << 8);
            if (r == 0.0 && i >= l) continue;
            D(l) -= p; E(l) = g; E(m) = 0.0;
        }
    }
    if (failed) { atomicExch(err_flag, 1); nh = 0; }
    tb.nhdr[2 * ik] = nh; tb.nhdr[2 * ik + 1] = failed ? 0 : pos;
    // ascending order (stable): rank of eigenvalue j, sorted copy
    double *Eo = tb.E + (size_t)ik * ne;
    unsigned char *rk = tb.rank + (size_t)ik * ne;
    for (int j = 0; j < n; ++j) {
        const double ej = D(j);
        int rnk = 0;
        for (int i = 0; i < n; ++i) { const double ei = D(i); rnk += (ei < ej || (ei == ej && i < j)) ? 1 : 0; }
        Eo[rnk] = ej; rk[j] = (unsigned char)rnk;
    }
#undef D
#undef E
}

// ================================================================================================ replay of the rotation log
// One thread per ROW of Z, the row in registers; the NE threads of a k-point (a "group": NE/32 warps) consume the same log
// strictly in order, so it is STREAMED through shared memory: a ring of kLogSlots chunks of kLogChunk rotations per group,
// filled by TMA bulk copies that lane 0 of the group's first warp issues, full/empty mbarriers per slot.  The rotations are
// then broadcast LDS.128 (tens of cycles) instead of dependent global loads (an exposed L2/HBM latency per 8 rotations was
// what bounded the first version: 78 ms per 110 K matrices of order 64, profiles/r01r).
constexpr int kLogChunk = 256, kLogSlots = 4;         // 4 x 256 x 16 B = 16 KB per k-point in flight

template <int NE>
__global__ void __launch_bounds__(128, 3)
zrot_kernel(const int n, const long long nk, const TriBuf tb) {
    constexpr int GROUPS = 128 / NE, GW = NE / 32;    // matrices per CTA, warps per matrix
    constexpr int RING = kLogChunk * kLogSlots;
    extern __shared__ __align__(128) unsigned char zsm[];       // [GROUPS][RING] double2, then [GROUPS][2][kLogSlots] mbarriers
    const int tid = threadIdx.x, lane = tid & 31;
    const int grp = tid / NE, row = tid % NE;
    const bool producer = (row == 0);                 // lane 0 of the group's first warp
    double2 *ring = reinterpret_cast<double2 *>(zsm) + (size_t)grp * RING;
    uint64_t *full = reinterpret_cast<uint64_t *>(zsm + sizeof(double2) * GROUPS * RING) + (size_t)grp * 2 * kLogSlots;
    uint64_t *empty = full + kLogSlots;
    if (producer) {
        for (int s = 0; s < kLogSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GW); }
        mbar_fence_init();
    }
    __syncthreads();
    // chunk counters run on across the k-points of this group, so the barrier phases stay consistent
    unsigned acquired = 0, released = 0, issued = 0;  // chunks waited for / handed back / requested (producer only)
    for (long long ik = (long long)blockIdx.x * GROUPS + grp; ik < nk; ik += (long long)gridDim.x * GROUPS) {
        double z[NE];
#pragma unroll
        for (int j = 0; j < NE; ++j) z[j] = (j == row) ? 1.0 : 0.0;
        const unsigned int *hdr = tb.hdr + (size_t)ik * tb.hdr_cap;
        const double2 *lg = tb.log + (size_t)ik * tb.log_cap;
        const int nh = tb.nhdr[2 * ik], nlog = tb.nhdr[2 * ik + 1];
        const unsigned nchunks = (unsigned)((nlog + kLogChunk - 1) / kLogChunk);
        const unsigned base = acquired;               // chunk counter value of this matrix's chunk 0 (== released == issued)
        const unsigned ring0 = base * kLogChunk;      // ring position of this matrix's log position 0
        auto issue_upto = [&](unsigned want) {        // producer: request chunks [issued, min(want, base + nchunks))
            while (issued < want && issued - base < nchunks) {
                const unsigned c = issued - base, slot = issued % kLogSlots;
                if (issued >= kLogSlots) mbar_wait(&empty[slot], ((issued / kLogSlots) + 1) & 1);
                const int cnt = (nlog - (int)c * kLogChunk < kLogChunk) ? nlog - (int)c * kLogChunk : kLogChunk;
                mbar_expect_tx(&full[slot], (uint32_t)cnt * 16u);
                tma_bulk_g2s(ring + (size_t)slot * kLogChunk, lg + (size_t)c * kLogChunk, (uint32_t)cnt * 16u, &full[slot]);
                ++issued;
            }
        };
        if (producer) issue_upto(base + kLogSlots);
        int pos0 = 0;
        for (int h = 0; h < nh; ++h) {
            const unsigned int w = __ldg(hdr + h);
            const int l = (int)(w & 0xffu), m = (int)(w >> 8);
            // the chase reads log positions [pos0, pos0 + m - l): make sure their chunks have landed
            const unsigned need = base + (unsigned)((pos0 + (m - l) - 1) / kLogChunk) + 1;
            while (acquired < need) { mbar_wait(&full[acquired % kLogSlots], (acquired / kLogSlots) & 1); ++acquired; }
            const int p_hi = pos0 + (m - 1);          // rotation of plane i sits at log position p_hi - i
#pragma unroll
            for (int blk = NE / 8 - 1; blk >= 0; --blk) {
                if (8 * blk + 7 <= NE - 2 && 8 * blk >= l && 8 * blk + 7 < m) {
                    // block entirely inside the chase (the common case of the long early chases): no clamps, no predicates —
                    // the replay is issue-bound (ncu r03v: issue 66 %, FP64 pipe 27 %), the index arithmetic is what costs
                    const unsigned r7 = ring0 + (unsigned)(p_hi - (8 * blk + 7));       // ring position of plane 8 blk + 7
                    double2 csb[8];
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) csb[ii] = ring[(r7 + (unsigned)(7 - ii)) % RING];
#pragma unroll
                    for (int ii = 7; ii >= 0; --ii) {
                        const int i = 8 * blk + ii;
                        const double2 cs = csb[ii];
                        const double f = z[i + 1 <= NE - 1 ? i + 1 : NE - 1], zi = z[i];
                        z[i + 1 <= NE - 1 ? i + 1 : NE - 1] = fma(cs.y, zi, cs.x * f);
                        z[i] = fma(-cs.y, f, cs.x * zi);                                // one dependent FMA per plane on the chain through f
                    }
                } else if (8 * blk < m && 8 * blk + 7 >= l) {
                    double2 csb[8];
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) {
                        int ic = 8 * blk + ii;
                        ic = ic < l ? l : (ic > m - 1 ? m - 1 : ic);
                        csb[ii] = ring[(ring0 + (unsigned)(p_hi - ic)) % RING];
                    }
#pragma unroll
                    for (int ii = 7; ii >= 0; --ii) {
                        const int i = 8 * blk + ii;
                        if (i <= NE - 2) {
                            if (i < m && i >= l) {
                                const double2 cs = csb[ii];
                                const double f = z[i + 1 <= NE - 1 ? i + 1 : NE - 1];
                                const double zi = z[i];
                                z[i + 1 <= NE - 1 ? i + 1 : NE - 1] = fma(cs.y, zi, cs.x * f);
                                z[i] = fma(-cs.y, f, cs.x * zi);
                            }
                        }
                    }
                }
            }
            pos0 += m - l;
            // hand back the chunks that lie completely behind pos0
            while ((int)(released - base + 1) * kLogChunk <= pos0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[released % kLogSlots]);
                ++released;
                if (producer) issue_upto(released + kLogSlots);
            }
        }
        // the tail chunk(s) of this matrix
        while (released - base < nchunks) {
            if (acquired <= released) { mbar_wait(&full[acquired % kLogSlots], (acquired / kLogSlots) & 1); ++acquired; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[released % kLogSlots]);
            ++released;
        }
        if (producer) issued = released;              // nothing beyond this matrix's chunks was requested
        double *Zo = tb.Z + (size_t)ik * NE * NE;
#pragma unroll
        for (int j = 0; j < NE; ++j) Zo[j * NE + row] = z[j];
    }
    (void)n;
}

}  // namespace

// ================================================================================================= launchers
int tri_padded_size(int nw) { return nw <= 32 ? 32 : 64; }
size_t tri_log_cap(int ne) { return (size_t)2 * ne * ne; }

cudaError_t launch_tridiag(int nw, long long nk, const double2 *Hk, const TriBuf &tb, int sm_count, cudaStream_t st) {
    const int grid = (int)std::min<long long>(nk < 1 ? 1 : nk, (long long)sm_count * 2);      // two CTAs of 256 threads per SM (110 / 128 registers)
    if (tb.ne == 32) tridiag_kernel<32, 8><<<grid, 256, 0, st>>>(nw, nk, Hk, tb);
    else tridiag_kernel<64, 4><<<grid, 256, 0, st>>>(nw, nk, Hk, tb);
    return cudaGetLastError();
}

cudaError_t launch_tql(int nw, long long nk, const TriBuf &tb, int *err_flag, cudaStream_t st) {
    const int grid = (int)((nk + kTqlThreads - 1) / kTqlThreads);
    const size_t smem = sizeof(double) * 2 * (size_t)tb.ne * kTqlThreads;
    tql_kernel<<<grid < 1 ? 1 : grid, kTqlThreads, smem, st>>>(nw, nk, tb, err_flag);
    return cudaGetLastError();
}

cudaError_t launch_zrot(int nw, long long nk, const TriBuf &tb, int sm_count, cudaStream_t st) {
    const int groups = 128 / tb.ne;
    long long grid = (nk + groups - 1) / groups;
    grid = std::max<long long>(1, std::min<long long>(grid, (long long)sm_count * (tb.ne == 32 ? 3 : 3)));
    const size_t smem = sizeof(double2) * groups * kLogChunk * kLogSlots + sizeof(uint64_t) * groups * 2 * kLogSlots;
    cudaError_t e;
    if (tb.ne == 32) {
        e = cudaFuncSetAttribute(zrot_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        zrot_kernel<32><<<(int)grid, 128, smem, st>>>(nw, nk, tb);
    } else {
        e = cudaFuncSetAttribute(zrot_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        zrot_kernel<64><<<(int)grid, 128, smem, st>>>(nw, nk, tb);
    }
    return cudaGetLastError();
}

}  // namespace mepb
