// mepb_ctx.cuh — host-side state of one mepb_ctx, shared by the translation units that implement the C ABI
// (mepb_api.cu: context, schedule, finalize; mepb_comm.cu: NCCL and the single-process multi-GPU driver).
#pragma once
#include <string>
#include <vector>

#include "mepb_internal.cuh"

namespace mepb {

int fail(int code, const std::string &msg);      // sets the thread's mepb_last_error text, returns code
const std::string &last_error_text();
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return mepb::fail(MEPB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));        \
    } while (0)

struct PackedLayout {  // identical to oracle/cpu_ref.c make_layout (documented in include/mepb.h)
    long long total, ne, mep_ic, mep_lc, kmep_ic, kmep_lc, ahc, velo, ohc, photo;
    long long mep_cs, kmep_cs, opt_h, opt_ah, gyro_c, gyro_d, gyro_dw;      // position-operator / gyro blocks (-1: absent)
};
inline PackedLayout make_layout(const mepb_params &p, bool posop) {
    PackedLayout L;
    long long o = 2;
    L.ne = o; o += p.n_eF;
    L.mep_ic = L.mep_lc = L.kmep_ic = L.kmep_lc = L.ahc = L.velo = L.ohc = L.photo = -1;
    if (p.do_mep) { L.mep_ic = o; o += 9LL * p.valence_bands; L.mep_lc = o; o += 9LL * p.valence_bands; }
    if (p.do_kubo) { L.kmep_ic = o; o += 9LL * p.n_eF; L.kmep_lc = o; o += 9LL * p.n_eF; }
    if (p.do_ahc) { L.ahc = o; o += 9LL * p.n_eF; L.velo = o; o += 18LL * p.n_hw * p.n_eF; L.ohc = o; o += 18LL * p.n_hw * p.n_eF; }
    if (p.do_opt) { L.photo = o; o += 54LL * p.n_hw * p.n_eF; }
    L.mep_cs = L.kmep_cs = L.opt_h = L.opt_ah = L.gyro_c = L.gyro_d = L.gyro_dw = -1;
    if (posop && p.do_mep) { L.mep_cs = o; o += p.valence_bands; }            // the scalar cs(i,i,n0) per band
    if (posop && p.do_kubo) { L.kmep_cs = o; o += p.n_eF; }
    if (posop && p.do_opt) { L.opt_h = o; o += 18LL * p.n_hw * p.n_eF; L.opt_ah = o; o += 18LL * p.n_hw * p.n_eF; }
    if (p.do_gyro) { L.gyro_c = o; o += 18LL * p.n_eF; L.gyro_d = o; o += 18LL * p.n_eF; L.gyro_dw = o; o += 18LL * p.n_hw * p.n_eF; }
    L.total = o;
    return L;
}

// compact per-eF accumulator block on the device: [static NS | sweep n_hw*36 | photo n_hw*54 | position-operator row]
struct CompactShape {
    int ns, sweep, photo, pos, per_ef;
};

struct StreamBuf {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    double *items_ahc = nullptr; int *cnt_ahc = nullptr;
    double *items_photo = nullptr; int *cnt_photo = nullptr;
    double *block_static = nullptr, *block_sweep = nullptr, *block_photo = nullptr;
    double *compact = nullptr;   // [n_eF][per_ef]
    double *eig_E = nullptr; double2 *eig_U = nullptr;   // [chunk][nw], [chunk][nw*nw] from the thread-per-k eigensolver
    double2 *V_pad = nullptr;                             // [chunk][3 nw (nw + 1)] H-gauge velocities, bands_sub_kernel -> mep_sums_kernel
    double2 *Hk = nullptr, *Hka = nullptr, *Vscr = nullptr;   // general path: [chunk][nw^2], [chunk][3][nw^2], [grid][3][nw^2]
    TriBuf tb{};                                                // general path: Householder + QL intermediates (mepb_tri.cu)
    double2 *G1 = nullptr, *G2 = nullptr;                       // general path, mesh mode: separable-transform tables
    // position-operator pipeline (mepb_posop.cu): W-gauge A_a [k][nw^2] and dA_a(b) [k][3][nw^2] per operand r_a, eigenpairs,
    // H-gauge V (user layout) / A [k][3][nw^2] / curvature pseudo-vectors [k][3][nw], scratch [grid][6][nw^2], block rows
    double2 *Aw[3] = {nullptr, nullptr, nullptr}, *dA[3] = {nullptr, nullptr, nullptr};
    double *pE = nullptr; double2 *pU = nullptr, *pV = nullptr, *pAH = nullptr, *pOmV = nullptr, *pScr = nullptr;
    double *prow = nullptr;
};

}  // namespace mepb

struct mepb_ctx {
    int device = 0, sm_count = 0;
    mepb_params p{};
    int nw = 0, nR = 0;
    bool has_tau = false;
    bool herm = true;      // H(R) Hermitian-closed: V_a Hermitian -> compact items / 18-accumulator sweep
    std::string herm_why;  // why the closure test failed (mepb_model_info)
    bool eig_thread = false;   // eigenpairs from the thread-per-k kernel (nw 4, 8)
    bool large = false;        // general path (mepb_large.cu)
    bool use_sub = false;      // NW-lanes-per-k band-sum kernel (mepb_sub.cu) in mesh mode
    bool tri = false;          // general path: Householder + QL eigensolver (default) instead of the in-kernel Jacobi
    bool pos = false;          // model carries a position operator r(R): Berry connection / curvature branch
    bool pipe2 = false;        // position-operator pipeline in use (pos, or do_gyro which needs V and the Kubo curvature per band)
    double2 *d_r = nullptr, *d_Hp_r[3] = {nullptr, nullptr, nullptr};   // r(R) (3,nw,nw,nR) and its packed element tiles per component
    int grid_pg = 0, grid_ps = 0, pos_tiles = 1, pos_row = 0;
    mepb::FourierPlan fp{};
    int *d_ridx = nullptr, *d_elem = nullptr; double2 *d_Hp = nullptr; double *d_Rfp = nullptr, *d_cRp = nullptr;
    double2 *d_Hka_int = nullptr;
    mepb::SepPlan sp{};              // mesh-separable Fourier transform (general path, mesh mode)
    int *d_map3 = nullptr, *d_ra = nullptr; double2 *d_P1 = nullptr, *d_P2 = nullptr, *d_P3 = nullptr; double *d_cw4 = nullptr;
    double Mmat[9];            // c_a(R) = sum_j Mmat[a + 3 j] R_j
    int *d_sp_ptr = nullptr, *d_sp_R = nullptr; double2 *d_sp_val = nullptr;   // sparse H(R) (warp-per-k Fourier sums)
    bool sparse = false; int sp_nnz = 0;
    int sp_up = 0;             // Hermitian-closed model: upper-triangle lane schedule appended to d_sp_ptr (mepb_api.cu build_sparse)
    int *d_ell_R = nullptr; double2 *d_ell_val = nullptr; double *d_ell_x = nullptr; int ell_n = 0, ell_nR = 0;   // mepb_eig.cu
    std::vector<double> h_Rf, h_tau;   // host copies (fractional R, centres) for rebuilding the sparse forms in update_model
    double a_latt[9];
    // device model
    double2 *d_H = nullptr; double *d_Rf = nullptr, *d_cR = nullptr, *d_tau = nullptr, *d_tM = nullptr, *d_hw = nullptr;
    std::vector<double> ef_lst, hw_lst;
    // schedule
    long long chunk = 0;
    int max_pairs = 0, max_photo = 0;
    int grid_interp = 0, grid_mep = 0, grid_sweep = 0, ns_sweep = 0, grid_photo = 0, ns_photo = 0;
    mepb::CompactShape cs{};
    mepb::PackedLayout L{};
    mepb::StreamBuf sb[2];
    double *d_packed = nullptr;
    int *d_err = nullptr;
    bool packed_from_host = false, have_data = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0;
    long long launches = 0;
    long long n_items_ahc = 0, n_items_photo = 0;
    bool counters_valid = false;   // the two work counters behind d_packed are newer than n_items_*
    // optional per-kernel-class timing (serialises the schedule onto one stream; used by bench.py for the roofline)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;          // pairs (start, stop)
    std::vector<int> prof_kind;
    double prof_ms[MEPB_PROF_KINDS] = {};
    long long prof_n[MEPB_PROF_KINDS] = {};
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;  // mepb_timer_start / stop
    // list-mode scratch
    double *d_kpt = nullptr, *d_e = nullptr; double2 *d_V = nullptr, *d_Hk = nullptr, *d_Hka = nullptr;
    long long list_cap = 0;
    // multi-GPU (mepb_comm.cu): NCCL communicator of this rank, stage times of the last step (ms; reset by mepb_reset)
    void *nccl_comm = nullptr; int comm_rank = 0, comm_size = 1;
    cudaEvent_t ar0 = nullptr, ar1 = nullptr;
    double t_acc_ms = 0.0, t_ar_ms = 0.0, t_fin_ms = 0.0;
    // mepb_band_sums staging (supplied eigenpairs)
    double *d_bs_e = nullptr; double2 *d_bs_V = nullptr; long long bs_cap = 0;
};


namespace mepb {
int pack_now(mepb_ctx *c);
void comm_release(mepb_ctx *c);                  // destroys the context's NCCL communicator, if any (mepb_comm.cu)                       // compact accumulators of both streams -> d_packed (device), on stream 0
}  // namespace mepb
