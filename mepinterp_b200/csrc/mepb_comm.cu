// mepb_comm.cu — the k-mesh exchange step inside the library: NCCL all-reduce of the packed partial sums, and a
// single-process driver over several B200s.
//
// Replaces src/mpi_comm.f90 for this path: the reference sums every accumulator with one MPI_REDUCE per tensor after the
// k-loop (core.f90:467-526) and broadcasts the chosen Fermi index (core.f90:409).  Here ONE ncclAllReduce(sum, double) of
// the packed buffer (a few KB — latency only, over NVLink 5 / NVSwitch) runs on the library's own launch stream right
// behind the packing kernel, with no host round trip in between, and every rank finalizes redundantly.
//   * one process per GPU (bench.py under torchrun, an MPI host): mepb_comm_unique_id on rank 0, broadcast the 128 bytes
//     by any means, mepb_comm_init_rank on every rank, then per step mepb_accumulate_strided + mepb_allreduce_partials
//     + mepb_finalize;
//   * one process, all GPUs (the C++ / Fortran driver): mepb_create_multi + mepb_multi_run_mesh (one host thread per
//     device, ncclCommInitAll).
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring a copy already loaded into the process, e.g. by torch), so
// libmepb.so keeps loading on boxes without NCCL and single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mepb_ctx.cuh"

using namespace mepb;

namespace {

struct NcclApi {
    bool tried = false, ok = false;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (api.tried) return api;
    api.tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // a copy the host process already loaded
    if (!h) if (const char *p = getenv("MEPB_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = std::string("libnccl.so.2 not found: ") + dlerror(); return api; }
#define MEPB_SYM(field, name)                                                             \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));                    \
    if (!api.field) { api.err = std::string("NCCL symbol missing: ") + name; return api; }
    MEPB_SYM(GetUniqueId, "ncclGetUniqueId");
    MEPB_SYM(CommInitRank, "ncclCommInitRank");
    MEPB_SYM(CommInitAll, "ncclCommInitAll");
    MEPB_SYM(CommDestroy, "ncclCommDestroy");
    MEPB_SYM(AllReduce, "ncclAllReduce");
    MEPB_SYM(GroupStart, "ncclGroupStart");
    MEPB_SYM(GroupEnd, "ncclGroupEnd");
    MEPB_SYM(GetErrorString, "ncclGetErrorString");
    MEPB_SYM(GetVersion, "ncclGetVersion");
#undef MEPB_SYM
    api.ok = true;
    return api;
}

#define NCCL_TRY(expr)                                                                                         \
    do {                                                                                                       \
        ncclResult_t r_ = (expr);                                                                              \
        if (r_ != ncclSuccess) return fail(MEPB_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(r_)); \
    } while (0)

int need_nccl() {
    NcclApi &a = nccl();
    if (!a.ok) return fail(MEPB_ERR_NCCL, a.err);
    return MEPB_OK;
}

int ensure_events(mepb_ctx *c) {
    if (!c->ar0) { CUDA_TRY(cudaEventCreate(&c->ar0)); CUDA_TRY(cudaEventCreate(&c->ar1)); }
    return MEPB_OK;
}

}  // namespace

void mepb::comm_release(mepb_ctx *c) {
    if (c->nccl_comm && nccl().ok) { cudaSetDevice(c->device); nccl().CommDestroy((ncclComm_t)c->nccl_comm); }
    c->nccl_comm = nullptr; c->comm_size = 1; c->comm_rank = 0;
    if (c->ar0) { cudaEventDestroy(c->ar0); cudaEventDestroy(c->ar1); c->ar0 = c->ar1 = nullptr; }
}

struct mepb_multi {
    std::vector<mepb_ctx *> ctx;
    int64_t block = 0;       // k-points per round-robin block (mepb_balanced_block)
};

extern "C" {

int mepb_comm_unique_id(void *id, int nbytes) {
    if (!id || nbytes < (int)sizeof(ncclUniqueId)) return fail(MEPB_ERR_ARG, "unique id buffer must hold 128 bytes");
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId u;
    NCCL_TRY(nccl().GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return MEPB_OK;
}

int mepb_comm_init_rank(mepb_ctx *c, int nranks, int rank, const void *id) {
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(MEPB_ERR_ARG, "bad argument");
    int rc = need_nccl();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(c->device));
    comm_release(c);
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    ncclComm_t comm;
    NCCL_TRY(nccl().CommInitRank(&comm, nranks, u, rank));
    c->nccl_comm = comm; c->comm_size = nranks; c->comm_rank = rank;
    return ensure_events(c);
}

int mepb_comm_size(const mepb_ctx *c) { return c ? c->comm_size : 0; }
int mepb_comm_rank(const mepb_ctx *c) { return c ? c->comm_rank : 0; }

// pack + in-place ncclAllReduce of the packed buffer on the context's launch stream; mepb_finalize / mepb_get_partials
// afterwards see the sum over all ranks.  A context without a communicator (one rank) only packs.
int mepb_allreduce_partials(mepb_ctx *c) {
    if (!c) return fail(MEPB_ERR_ARG, "null ctx");
    if (!c->have_data) return fail(MEPB_ERR_STATE, "allreduce before accumulate");
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->packed_from_host) { int rc = pack_now(c); if (rc) return rc; }
    c->packed_from_host = true;
    if (c->nccl_comm && c->comm_size > 1) {
        int rc = ensure_events(c);
        if (rc) return rc;
        cudaStream_t st = c->sb[0].st;
        CUDA_TRY(cudaEventRecord(c->ar0, st));
        NCCL_TRY(nccl().AllReduce(c->d_packed, c->d_packed, (size_t)c->L.total, ncclDouble, ncclSum, (ncclComm_t)c->nccl_comm, st));
        CUDA_TRY(cudaEventRecord(c->ar1, st));
        CUDA_TRY(cudaEventSynchronize(c->ar1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->ar0, c->ar1));
        c->t_ar_ms += ms;
    }
    return MEPB_OK;
}

// ------------------------------------------------------------------------------------------------ one process, several GPUs
int mepb_create_multi(const mepb_model *model, const mepb_params *params, int n_devices, const int *devices, mepb_multi **out) {
    if (!out || n_devices < 1) return fail(MEPB_ERR_ARG, "bad argument");
    *out = nullptr;
    mepb_multi *mm = new mepb_multi();
    std::vector<int> dev(n_devices);
    for (int i = 0; i < n_devices; ++i) dev[i] = devices ? devices[i] : i;
    for (int i = 0; i < n_devices; ++i) {
        mepb_ctx *c = nullptr;
        const int rc = mepb_create(model, params, dev[i], &c);
        if (rc) { for (mepb_ctx *x : mm->ctx) mepb_destroy(x); delete mm; return rc; }
        mm->ctx.push_back(c);
    }
    if (n_devices > 1) {
        int rc = need_nccl();
        std::vector<ncclComm_t> comms(n_devices);
        if (!rc) {
            ncclResult_t r = nccl().CommInitAll(comms.data(), n_devices, dev.data());
            if (r != ncclSuccess) rc = fail(MEPB_ERR_NCCL, std::string("ncclCommInitAll: ") + nccl().GetErrorString(r));
        }
        if (rc) { for (mepb_ctx *x : mm->ctx) mepb_destroy(x); delete mm; return rc; }
        for (int i = 0; i < n_devices; ++i) {
            mm->ctx[i]->nccl_comm = comms[i]; mm->ctx[i]->comm_size = n_devices; mm->ctx[i]->comm_rank = i;
            cudaSetDevice(dev[i]);
            ensure_events(mm->ctx[i]);
        }
    }
    const mepb_params &p = *params;
    const long long nk = (long long)p.mp_grid[0] * p.mp_grid[1] * p.mp_grid[2];
    long long chunk = mm->ctx[0]->chunk;
    for (mepb_ctx *x : mm->ctx) chunk = std::min(chunk, x->chunk);
    for (mepb_ctx *x : mm->ctx) x->chunk = chunk;              // same wave size everywhere: identical blocks on every device
    mm->block = mepb_balanced_block(mm->ctx[0], nk, n_devices);
    *out = mm;
    return MEPB_OK;
}

int mepb_multi_size(const mepb_multi *mm) { return mm ? (int)mm->ctx.size() : 0; }
mepb_ctx *mepb_multi_ctx(mepb_multi *mm, int i) { return (mm && i >= 0 && i < (int)mm->ctx.size()) ? mm->ctx[i] : nullptr; }

// core_worker over all devices: every device accumulates its round-robin blocks of the mesh (one host thread per device, so the
// launch schedules run concurrently), one NCCL all-reduce, finalize on the first device.
int mepb_multi_run_mesh(mepb_multi *mm, mepb_result *out) {
    if (!mm || !out) return fail(MEPB_ERR_ARG, "null argument");
    const int n = (int)mm->ctx.size();
    const mepb_params &p = mm->ctx[0]->p;
    const long long nk = (long long)p.mp_grid[0] * p.mp_grid[1] * p.mp_grid[2];
    std::vector<int> rcs(n, MEPB_OK);
    std::vector<std::string> msgs(n);
    auto work = [&](int i) {
        mepb_ctx *c = mm->ctx[i];
        int rc = mepb_reset(c);
        if (!rc) rc = mepb_accumulate_strided(c, 0, nk, mm->block, i, n);
        if (rc) { rcs[i] = rc; msgs[i] = last_error_text(); }
    };
    {
        std::vector<std::thread> th;
        for (int i = 1; i < n; ++i) th.emplace_back(work, i);
        work(0);
        for (auto &t : th) t.join();
    }
    for (int i = 0; i < n; ++i) if (rcs[i]) return fail(rcs[i], "device " + std::to_string(mm->ctx[i]->device) + ": " + msgs[i]);
    // the collective is entered only when every device finished its slab without error (a failed rank would hang the others)
    auto reduce = [&](int i) {
        const int rc = mepb_allreduce_partials(mm->ctx[i]);
        if (rc) { rcs[i] = rc; msgs[i] = last_error_text(); }
    };
    {
        std::vector<std::thread> th;
        for (int i = 1; i < n; ++i) th.emplace_back(reduce, i);
        reduce(0);
        for (auto &t : th) t.join();
    }
    for (int i = 0; i < n; ++i) if (rcs[i]) return fail(rcs[i], "device " + std::to_string(mm->ctx[i]->device) + ": " + msgs[i]);
    return mepb_finalize(mm->ctx[0], out);
}

int mepb_multi_update_model(mepb_multi *mm, const mepb_model *model) {
    if (!mm) return fail(MEPB_ERR_ARG, "null argument");
    for (mepb_ctx *c : mm->ctx) { const int rc = mepb_update_model(c, model); if (rc) return rc; }
    return MEPB_OK;
}

void mepb_multi_destroy(mepb_multi *mm) {
    if (!mm) return;
    for (mepb_ctx *c : mm->ctx) mepb_destroy(c);
    delete mm;
}

}  // extern "C"
