// dist.cu -- the multi-GPU layer of the C ABI: particles shard, the mesh and the resident snapshots are
// replicated, and the one exchange of the path -- recorded trajectories and end points back to one owner, in caller order
// -- runs over NCCL (NVLink / NVSwitch).  Two forms over one implementation:
//   mops_dist_*   one process per GPU (torchrun, MPI ...): every process owns one mops_ctx and joins a communicator
//                 built from a caller-broadcast ncclUniqueId;
//   mops_multi_*  one process, N GPUs: N contexts + one worker thread per device behind a single handle; this is what the
//                 C++ drop-in (MOPS_RunStreamLine / MOPS_RunPathLine) uses when more than one device is selected.
// The reference has nothing to restate here: its only multi-rank code is a serial loop over MPI ranks in the CLI
// (CLI/main.cpp:58-66, 276-284).  Caller-order reassembly follows TrajectoryCommon.h:47,124 (lineID = input index).
// NCCL is loaded at run time (dlopen of libnccl.so.2: the copy PyTorch has already loaded in a torch process, the system
// one otherwise), so single-GPU users of libmops_b200.so carry no NCCL dependency.
#include "../../include/mops_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        auto sym = [&](const char* n) { return dlsym(api.lib, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Send && api.Recv && api.Broadcast &&
                 api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api;
}

struct DBuf { // grow-only device scratch
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return MOPS_E_NOMEM; }
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// dst[idx[row]] = src[row]: rows of rw 4-byte words, one thread per word (reads coalesced, each row written as one segment)
__global__ void k_scatter_rows(const unsigned* __restrict__ src, const int* __restrict__ idx, unsigned* __restrict__ dst, long long n_rows, int rw)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * rw) return;
    const long long row = t / rw;
    const int j = (int)(t - row * rw);
    dst[(long long)idx[row] * rw + j] = src[t];
}
// dst[row] = src[idx[row]]
__global__ void k_gather_rows(const unsigned* __restrict__ src, const int* __restrict__ idx, unsigned* __restrict__ dst, long long n_rows, int rw)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * rw) return;
    const long long row = t / rw;
    const int j = (int)(t - row * rw);
    dst[t] = src[(long long)idx[row] * rw + j];
}
__global__ void k_iota32(int* __restrict__ a, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}

inline int nblk(long long n, int bs) { return (int)((n + bs - 1) / bs); }

} // namespace

struct mops_dist {
    mops_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    bool own_comm = true;
    int rank = 0, world = 1, device = 0;
    DBuf st_idx, st_rows; // root-side staging of a gather
    cudaStream_t stream = nullptr; // null: the context's stream
    std::string err;
};

namespace {
int dfail(mops_dist* d, int code, const char* what, const char* detail)
{
    if (d) d->err = std::string(what) + ": " + (detail ? detail : "");
    return code;
}
#define DCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) return dfail(d, MOPS_E_CUDA, #call, cudaGetErrorString(e__));              \
    } while (0)
#define NCK(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess) return dfail(d, MOPS_E_CUDA, #call, nccl().GetErrorString(r__));           \
    } while (0)

struct GArr { // one array of a gather: rows of `words` 4-byte words; src on every rank, dst on the root
    const void* src;
    void* dst;
    int words;
};

// the gather itself: rows of every rank -> root, scattered to their caller index.  counts[r] = rows of rank r.
int gather_impl(mops_dist* d, int root, int64_t n_local, const int64_t* counts, const int32_t* d_index, const GArr* arr, int n_arr,
                cudaStream_t st)
{
    NcclApi& N = nccl();
    DCK(cudaSetDevice(d->device));
    if (d->rank != root) {
        if (n_local == 0) return MOPS_OK;
        NCK(N.GroupStart());
        NCK(N.Send(d_index, (size_t)n_local, ncclInt32, root, d->comm, st));
        for (int a = 0; a < n_arr; ++a)
            if (arr[a].src) NCK(N.Send(arr[a].src, (size_t)n_local * arr[a].words, ncclInt32, root, d->comm, st));
        NCK(N.GroupEnd());
        return MOPS_OK;
    }
    int64_t remote = 0;
    size_t words_total = 0;
    for (int r = 0; r < d->world; ++r)
        if (r != root) remote += counts[r];
    for (int a = 0; a < n_arr; ++a)
        if (arr[a].src) words_total += (size_t)arr[a].words;
    if (remote > 0) {
        if (d->st_idx.ensure((size_t)remote * 4)) return dfail(d, MOPS_E_NOMEM, "gather staging", "index");
        if (d->st_rows.ensure((size_t)remote * words_total * 4)) return dfail(d, MOPS_E_NOMEM, "gather staging", "rows");
        int* s_idx = (int*)d->st_idx.p;
        NCK(N.GroupStart());
        int64_t off = 0;
        for (int r = 0; r < d->world; ++r) {
            if (r == root || counts[r] == 0) continue;
            const size_t c = (size_t)counts[r];
            NCK(N.Recv(s_idx + off, c, ncclInt32, r, d->comm, st));
            size_t base = 0; // array a of all remote ranks is one contiguous staging block of remote * words
            for (int a = 0; a < n_arr; ++a) {
                if (!arr[a].src) continue;
                unsigned* blk = (unsigned*)d->st_rows.p + base;
                NCK(N.Recv(blk + (size_t)off * arr[a].words, c * arr[a].words, ncclInt32, r, d->comm, st));
                base += (size_t)remote * arr[a].words;
            }
            off += (int64_t)c;
        }
        NCK(N.GroupEnd());
    }
    // caller order: the root's own rows straight from its buffers, the received ones from the staging
    size_t base = 0;
    for (int a = 0; a < n_arr; ++a) {
        if (!arr[a].src) continue;
        const int w = arr[a].words;
        if (n_local) k_scatter_rows<<<nblk(n_local * w, 256), 256, 0, st>>>((const unsigned*)arr[a].src, d_index, (unsigned*)arr[a].dst, n_local, w);
        if (remote) k_scatter_rows<<<nblk(remote * w, 256), 256, 0, st>>>((const unsigned*)d->st_rows.p + base, (const int*)d->st_idx.p,
                                                                        (unsigned*)arr[a].dst, remote, w);
        base += (size_t)remote * w;
    }
    DCK(cudaGetLastError());
    return MOPS_OK;
}
} // namespace

extern "C" {

int mops_dist_unique_id(void* id128)
{
    if (!id128) return MOPS_E_INVALID;
    NcclApi& N = nccl();
    if (!N.ok) return MOPS_E_STATE;
    ncclUniqueId id;
    if (N.GetUniqueId(&id) != ncclSuccess) return MOPS_E_CUDA;
    std::memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return MOPS_OK;
}

int mops_dist_create(mops_dist** out, mops_ctx* ctx, int32_t rank, int32_t world, const void* id128)
{
    if (!out) return MOPS_E_INVALID;
    *out = nullptr;
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return MOPS_E_INVALID;
    NcclApi& N = nccl();
    if (!N.ok) return MOPS_E_STATE; // libnccl.so.2 not found: no multi-GPU
    mops_dist* d = new mops_dist();
    d->ctx = ctx; d->rank = rank; d->world = world; d->device = mops_get_device(ctx);
    if (cudaSetDevice(d->device) != cudaSuccess) { delete d; return MOPS_E_CUDA; }
    ncclUniqueId id;
    std::memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    if (N.CommInitRank(&d->comm, world, id, rank) != ncclSuccess) { delete d; return MOPS_E_CUDA; }
    *out = d;
    return MOPS_OK;
}

void mops_dist_destroy(mops_dist* d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
    if (d->comm && d->own_comm) nccl().CommDestroy(d->comm);
    d->st_idx.release(); d->st_rows.release();
    delete d;
}

const char* mops_dist_last_error(const mops_dist* d) { return d ? d->err.c_str() : "null handle"; }

void mops_shard_bounds(int64_t n_total, int32_t rank, int32_t world, int64_t* lo, int64_t* hi)
{
    // equal contiguous blocks of the key-sorted seed set; the first (n_total % world) blocks hold one more
    const int64_t base = n_total / world, extra = n_total % world;
    const int64_t l = rank * base + std::min<int64_t>(rank, extra);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (rank < extra ? 1 : 0);
}

int mops_dist_gather_traj(mops_dist* d, int32_t root, int64_t n_local, const int64_t* counts, const int32_t* index, int32_t each,
                          const double* pos, const double* vel, const double* xyz, const float* depth, int64_t n_total, double* out_pos,
                          double* out_vel, double* out_xyz, float* out_depth)
{
    if (!d || !counts || root < 0 || root >= d->world || n_local < 0 || each < 0) return MOPS_E_INVALID;
    if (n_local != counts[d->rank]) return dfail(d, MOPS_E_INVALID, "mops_dist_gather_traj", "n_local != counts[rank]");
    if (n_local > 0 && !index) return dfail(d, MOPS_E_INVALID, "mops_dist_gather_traj", "null index");
    if (d->rank == root && ((pos && !out_pos) || (vel && !out_vel) || (xyz && !out_xyz) || (depth && !out_depth)))
        return dfail(d, MOPS_E_INVALID, "mops_dist_gather_traj", "null output on the root");
    const GArr arr[4] = {{pos, out_pos, each * 6}, {vel, out_vel, each * 6}, {xyz, out_xyz, 6}, {depth, out_depth, 1}};
    return gather_impl(d, root, n_local, counts, index, arr, 4, d->stream ? d->stream : (cudaStream_t)mops_get_stream(d->ctx));
}

int mops_dist_set_stream(mops_dist* d, void* cuda_stream)
{
    if (!d) return MOPS_E_INVALID;
    d->stream = (cudaStream_t)cuda_stream;
    return MOPS_OK;
}

} // extern "C"

// =========================================================================================================================
// one process, N GPUs
// =========================================================================================================================
struct mops_multi {
    int n = 0;
    std::vector<int> devices;
    std::vector<mops_ctx*> ctx;
    std::vector<mops_dist*> dist;
    std::string err;
    // per-device particle buffers of the current call (grow-only)
    struct Dev {
        DBuf xyz, depth, cell0, index, out_pos, out_vel, out_attr, status, steps, fcell;
    };
    std::vector<Dev> dev;
    // root-side (device 0) buffers: all seeds, sort scratch, caller-order results
    DBuf r_xyz, r_depth, r_cell0, r_key, r_key2, r_val, r_perm, r_tmp, r_sx, r_sd, r_sc, r_out_pos, r_out_vel, r_out_attr, r_status, r_steps,
        r_fcell;
};

namespace {

// run fn(i) for every device on its own host thread (contexts are single-threaded, devices are independent)
int for_devices(mops_multi* mm, const std::function<int(int)>& fn)
{
    std::vector<int> rc(mm->n, 0);
    std::vector<std::thread> th;
    for (int i = 1; i < mm->n; ++i) th.emplace_back([&, i] { rc[i] = fn(i); });
    rc[0] = fn(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < mm->n; ++i)
        if (rc[i]) {
            const char* e = mops_last_error(mm->ctx[i]);
            const char* de = mm->dist[i] ? mm->dist[i]->err.c_str() : "";
            mm->err = "device " + std::to_string(mm->devices[i]) + ": " + ((e && *e) ? e : de);
            return rc[i];
        }
    return 0;
}

#define MCK(call)                                                                                            \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) { mm->err = std::string(#call) + ": " + cudaGetErrorString(e__); return MOPS_E_CUDA; } \
    } while (0)

int multi_traj(mops_multi* mm, const mops_traj_cfg* cfg, int front, int back, const mops_traj_io* io, mops_traj_stats* stats, bool path)
{
    if (!mm || !cfg || !io) return MOPS_E_INVALID;
    if (cfg->mem != MOPS_MEM_HOST) { mm->err = "mops_multi_* trajectory calls take HOST-memory buffers (caller order)"; return MOPS_E_INVALID; }
    const int64_t n = io->n;
    if (n < 0 || n > 0x7fffffffLL) { mm->err = "particle count out of range"; return MOPS_E_INVALID; }
    if (n == 0) {
        if (stats) std::memset(stats, 0, sizeof(*stats));
        return MOPS_OK;
    }
    if (cfg->delta_t <= 0 || cfg->record_t <= 0 || cfg->duration <= 0) { mm->err = "invalid trajectory settings"; return MOPS_E_INVALID; }
    if (!io->xyz || !io->depth || !io->out_pos || !io->out_vel) { mm->err = "null particle buffers"; return MOPS_E_INVALID; }
    if (io->out_cell_log || io->out_min_edge) { mm->err = "per-step cell logs / edge distances are single-device diagnostics"; return MOPS_E_INVALID; }
    const int each = (int)(cfg->duration / cfg->record_t);
    if (each <= 0 || cfg->duration / cfg->delta_t <= 0) { mm->err = "invalid integration steps"; return MOPS_E_INVALID; }
    const int G = mm->n;
    mops_ctx* c0 = mm->ctx[0];
    cudaStream_t s0 = (cudaStream_t)mops_get_stream(c0);
    MCK(cudaSetDevice(mm->devices[0]));

    // ---- root: seeds up, processing-order key (Morton rank of the start cell), sort, sorted copies ------------------
    if (mm->r_xyz.ensure((size_t)n * 24) || mm->r_depth.ensure((size_t)n * 4) || mm->r_key.ensure((size_t)n * 4) ||
        mm->r_key2.ensure((size_t)n * 4) || mm->r_val.ensure((size_t)n * 4) || mm->r_perm.ensure((size_t)n * 4) ||
        mm->r_sx.ensure((size_t)n * 24) || mm->r_sd.ensure((size_t)n * 4) || (io->cell0 && (mm->r_cell0.ensure((size_t)n * 4) || mm->r_sc.ensure((size_t)n * 4)))) {
        mm->err = "out of device memory (root seed buffers)";
        return MOPS_E_NOMEM;
    }
    MCK(cudaMemcpyAsync(mm->r_xyz.p, io->xyz, (size_t)n * 24, cudaMemcpyHostToDevice, s0));
    MCK(cudaMemcpyAsync(mm->r_depth.p, io->depth, (size_t)n * 4, cudaMemcpyHostToDevice, s0));
    if (io->cell0) MCK(cudaMemcpyAsync(mm->r_cell0.p, io->cell0, (size_t)n * 4, cudaMemcpyHostToDevice, s0));
    int rc = mops_order_key(c0, MOPS_MEM_DEVICE, n, (const double*)mm->r_xyz.p, (int32_t*)mm->r_key.p);
    if (rc) { mm->err = mops_last_error(c0); return rc; }
    k_iota32<<<nblk(n, 256), 256, 0, s0>>>((int*)mm->r_val.p, n);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned*)mm->r_key.p, (unsigned*)mm->r_key2.p, (const int*)mm->r_val.p,
                                    (int*)mm->r_perm.p, (int)n, 0, 32, s0);
    if (mm->r_tmp.ensure(tmp_bytes)) { mm->err = "out of device memory (sort scratch)"; return MOPS_E_NOMEM; }
    MCK(cub::DeviceRadixSort::SortPairs(mm->r_tmp.p, tmp_bytes, (const unsigned*)mm->r_key.p, (unsigned*)mm->r_key2.p, (const int*)mm->r_val.p,
                                        (int*)mm->r_perm.p, (int)n, 0, 32, s0));
    const int* perm = (const int*)mm->r_perm.p; // perm[j] = caller index of the j-th particle along the curve
    k_gather_rows<<<nblk(n * 6, 256), 256, 0, s0>>>((const unsigned*)mm->r_xyz.p, perm, (unsigned*)mm->r_sx.p, n, 6);
    k_gather_rows<<<nblk(n, 256), 256, 0, s0>>>((const unsigned*)mm->r_depth.p, perm, (unsigned*)mm->r_sd.p, n, 1);
    if (io->cell0) k_gather_rows<<<nblk(n, 256), 256, 0, s0>>>((const unsigned*)mm->r_cell0.p, perm, (unsigned*)mm->r_sc.p, n, 1);
    MCK(cudaGetLastError());

    // ---- equal contiguous blocks of the sorted set, one per device -------------------------------------------------------
    std::vector<int64_t> lo(G), hi(G), counts(G);
    for (int g = 0; g < G; ++g) {
        mops_shard_bounds(n, g, G, &lo[g], &hi[g]);
        counts[g] = hi[g] - lo[g];
    }
    const bool want_attr = path && io->out_attr;
    const size_t out_bytes_total = (size_t)n * each * 24;
    if (mm->r_out_pos.ensure(out_bytes_total) || mm->r_out_vel.ensure(out_bytes_total) || (want_attr && mm->r_out_attr.ensure(out_bytes_total)) ||
        (io->out_status && mm->r_status.ensure((size_t)n * 4)) || (io->out_steps && mm->r_steps.ensure((size_t)n * 4)) ||
        (io->out_cell && mm->r_fcell.ensure((size_t)n * 4))) {
        mm->err = "out of device memory (root result buffers)";
        return MOPS_E_NOMEM;
    }
    std::vector<mops_traj_stats> st(G);
    NcclApi& N = nccl();
    rc = for_devices(mm, [&](int g) -> int {
        mops_multi::Dev& D = mm->dev[g];
        mops_dist* d = mm->dist[g];
        mops_ctx* c = mm->ctx[g];
        cudaStream_t sg = (cudaStream_t)mops_get_stream(c);
        const int64_t m = counts[g];
        DCK(cudaSetDevice(mm->devices[g]));
        const size_t ob = (size_t)std::max<int64_t>(m, 1) * each * 24;
        if (D.xyz.ensure((size_t)std::max<int64_t>(m, 1) * 24) || D.depth.ensure((size_t)std::max<int64_t>(m, 1) * 4) ||
            D.index.ensure((size_t)std::max<int64_t>(m, 1) * 4) || (io->cell0 && D.cell0.ensure((size_t)std::max<int64_t>(m, 1) * 4)) ||
            D.out_pos.ensure(ob) || D.out_vel.ensure(ob) || (want_attr && D.out_attr.ensure(ob)) ||
            (io->out_status && D.status.ensure((size_t)std::max<int64_t>(m, 1) * 4)) || (io->out_steps && D.steps.ensure((size_t)std::max<int64_t>(m, 1) * 4)) ||
            (io->out_cell && D.fcell.ensure((size_t)std::max<int64_t>(m, 1) * 4)))
            return dfail(d, MOPS_E_NOMEM, "mops_multi", "out of device memory (shard buffers)");
        // the shard's seeds: device 0 keeps its block with a local copy, the others receive theirs over NVLink
        if (g == 0) {
            DCK(cudaMemcpyAsync(D.xyz.p, (const char*)mm->r_sx.p + (size_t)lo[0] * 24, (size_t)m * 24, cudaMemcpyDeviceToDevice, sg));
            DCK(cudaMemcpyAsync(D.depth.p, (const char*)mm->r_sd.p + (size_t)lo[0] * 4, (size_t)m * 4, cudaMemcpyDeviceToDevice, sg));
            DCK(cudaMemcpyAsync(D.index.p, (const char*)mm->r_perm.p + (size_t)lo[0] * 4, (size_t)m * 4, cudaMemcpyDeviceToDevice, sg));
            if (io->cell0) DCK(cudaMemcpyAsync(D.cell0.p, (const char*)mm->r_sc.p + (size_t)lo[0] * 4, (size_t)m * 4, cudaMemcpyDeviceToDevice, sg));
            if (G > 1) NCK(N.GroupStart());
            for (int r = 1; r < G; ++r) {
                if (counts[r] == 0) continue;
                NCK(N.Send((const char*)mm->r_sx.p + (size_t)lo[r] * 24, (size_t)counts[r] * 6, ncclInt32, r, d->comm, sg));
                NCK(N.Send((const char*)mm->r_sd.p + (size_t)lo[r] * 4, (size_t)counts[r], ncclInt32, r, d->comm, sg));
                NCK(N.Send((const char*)mm->r_perm.p + (size_t)lo[r] * 4, (size_t)counts[r], ncclInt32, r, d->comm, sg));
                if (io->cell0) NCK(N.Send((const char*)mm->r_sc.p + (size_t)lo[r] * 4, (size_t)counts[r], ncclInt32, r, d->comm, sg));
            }
            if (G > 1) NCK(N.GroupEnd());
        } else if (m > 0) {
            NCK(N.GroupStart());
            NCK(N.Recv(D.xyz.p, (size_t)m * 6, ncclInt32, 0, d->comm, sg));
            NCK(N.Recv(D.depth.p, (size_t)m, ncclInt32, 0, d->comm, sg));
            NCK(N.Recv(D.index.p, (size_t)m, ncclInt32, 0, d->comm, sg));
            if (io->cell0) NCK(N.Recv(D.cell0.p, (size_t)m, ncclInt32, 0, d->comm, sg));
            NCK(N.GroupEnd());
        }
        // integrate the shard (device-memory form of the single-GPU call)
        std::memset(&st[g], 0, sizeof(st[g]));
        if (m > 0) {
            mops_traj_cfg cg = *cfg;
            cg.mem = MOPS_MEM_DEVICE;
            mops_traj_io ig;
            std::memset(&ig, 0, sizeof(ig));
            ig.n = m; ig.xyz = (double*)D.xyz.p; ig.depth = (float*)D.depth.p; ig.cell0 = io->cell0 ? (const int32_t*)D.cell0.p : nullptr;
            ig.out_pos = (double*)D.out_pos.p; ig.out_vel = (double*)D.out_vel.p; ig.out_attr = want_attr ? (double*)D.out_attr.p : nullptr;
            ig.out_status = io->out_status ? (int32_t*)D.status.p : nullptr;
            ig.out_steps = io->out_steps ? (int32_t*)D.steps.p : nullptr;
            ig.out_cell = io->out_cell ? (int32_t*)D.fcell.p : nullptr;
            const int r2 = path ? mops_pathline(c, &cg, front, back, &ig, &st[g]) : mops_streamline(c, &cg, front, &ig, &st[g]);
            if (r2) return r2;
        }
        // the path's one exchange: records + end points back to device 0, in caller order, over NCCL
        const GArr arr[8] = {{D.out_pos.p, mm->r_out_pos.p, each * 6}, {D.out_vel.p, mm->r_out_vel.p, each * 6},
                             {want_attr ? D.out_attr.p : nullptr, mm->r_out_attr.p, each * 6}, {D.xyz.p, mm->r_xyz.p, 6}, {D.depth.p, mm->r_depth.p, 1},
                             {io->out_status ? D.status.p : nullptr, mm->r_status.p, 1}, {io->out_steps ? D.steps.p : nullptr, mm->r_steps.p, 1},
                             {io->out_cell ? D.fcell.p : nullptr, mm->r_fcell.p, 1}};
        const int r3 = gather_impl(d, 0, m, counts.data(), (const int32_t*)D.index.p, arr, 8, sg);
        if (r3) return r3;
        DCK(cudaStreamSynchronize(sg));
        return 0;
    });
    if (rc) return rc;

    // ---- root: caller-order results down to the caller's buffers ------------------------------------------------------------
    MCK(cudaSetDevice(mm->devices[0]));
    MCK(cudaMemcpyAsync(io->xyz, mm->r_xyz.p, (size_t)n * 24, cudaMemcpyDeviceToHost, s0));
    MCK(cudaMemcpyAsync(io->depth, mm->r_depth.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s0));
    MCK(cudaMemcpyAsync(io->out_pos, mm->r_out_pos.p, out_bytes_total, cudaMemcpyDeviceToHost, s0));
    MCK(cudaMemcpyAsync(io->out_vel, mm->r_out_vel.p, out_bytes_total, cudaMemcpyDeviceToHost, s0));
    if (want_attr) MCK(cudaMemcpyAsync(io->out_attr, mm->r_out_attr.p, out_bytes_total, cudaMemcpyDeviceToHost, s0));
    if (io->out_status) MCK(cudaMemcpyAsync(io->out_status, mm->r_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s0));
    if (io->out_steps) MCK(cudaMemcpyAsync(io->out_steps, mm->r_steps.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s0));
    if (io->out_cell) MCK(cudaMemcpyAsync(io->out_cell, mm->r_fcell.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s0));
    MCK(cudaStreamSynchronize(s0));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (int g = 0; g < G; ++g) {
            stats->particle_steps += st[g].particle_steps;
            stats->alive_at_end += st[g].alive_at_end;
            stats->near_edge_particles += st[g].near_edge_particles;
            stats->above_surface_particles += st[g].above_surface_particles;
            stats->kernel_ms = std::max(stats->kernel_ms, st[g].kernel_ms);
            stats->locate_ms = std::max(stats->locate_ms, st[g].locate_ms);
            stats->total_ms = std::max(stats->total_ms, st[g].total_ms);
            stats->launches += st[g].launches;
        }
    }
    return MOPS_OK;
}

} // namespace

extern "C" {

int mops_multi_create(mops_multi** out, int32_t n_devices, const int32_t* devices)
{
    if (!out) return MOPS_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return MOPS_E_NODEVICE;
    if (n_devices <= 0) n_devices = ndev; // all devices of the box
    if (n_devices > ndev) return MOPS_E_NODEVICE;
    mops_multi* mm = new mops_multi();
    mm->n = n_devices;
    for (int i = 0; i < n_devices; ++i) mm->devices.push_back(devices ? devices[i] : i);
    mm->ctx.assign(n_devices, nullptr);
    mm->dist.assign(n_devices, nullptr);
    mm->dev.resize(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        const int rc = mops_create(&mm->ctx[i], mm->devices[i]);
        if (rc) { mops_multi_destroy(mm); return rc; }
    }
    if (n_devices > 1) {
        NcclApi& N = nccl();
        if (!N.ok) { mops_multi_destroy(mm); return MOPS_E_STATE; }
        std::vector<ncclComm_t> comms(n_devices);
        if (N.CommInitAll(comms.data(), n_devices, mm->devices.data()) != ncclSuccess) { mops_multi_destroy(mm); return MOPS_E_CUDA; }
        for (int i = 0; i < n_devices; ++i) {
            mops_dist* d = new mops_dist();
            d->ctx = mm->ctx[i]; d->comm = comms[i]; d->rank = i; d->world = n_devices; d->device = mm->devices[i];
            mm->dist[i] = d;
        }
    } else {
        mops_dist* d = new mops_dist(); // world of one: the gather degenerates to the local scatter
        d->ctx = mm->ctx[0]; d->comm = nullptr; d->own_comm = false; d->rank = 0; d->world = 1; d->device = mm->devices[0];
        mm->dist[0] = d;
    }
    *out = mm;
    return MOPS_OK;
}

void mops_multi_destroy(mops_multi* mm)
{
    if (!mm) return;
    for (int i = 0; i < mm->n; ++i) {
        cudaSetDevice(mm->devices[i]);
        cudaDeviceSynchronize();
        if (i < (int)mm->dev.size()) {
            mops_multi::Dev& D = mm->dev[i];
            for (DBuf* b : {&D.xyz, &D.depth, &D.cell0, &D.index, &D.out_pos, &D.out_vel, &D.out_attr, &D.status, &D.steps, &D.fcell}) b->release();
        }
        if (i == 0)
            for (DBuf* b : {&mm->r_xyz, &mm->r_depth, &mm->r_cell0, &mm->r_key, &mm->r_key2, &mm->r_val, &mm->r_perm, &mm->r_tmp, &mm->r_sx, &mm->r_sd,
                            &mm->r_sc, &mm->r_out_pos, &mm->r_out_vel, &mm->r_out_attr, &mm->r_status, &mm->r_steps, &mm->r_fcell})
                b->release();
        if (i < (int)mm->dist.size() && mm->dist[i]) mops_dist_destroy(mm->dist[i]);
        if (i < (int)mm->ctx.size() && mm->ctx[i]) mops_destroy(mm->ctx[i]);
    }
    delete mm;
}

const char* mops_multi_last_error(const mops_multi* mm) { return mm ? mm->err.c_str() : "null handle"; }
int32_t mops_multi_device_count(const mops_multi* mm) { return mm ? mm->n : 0; }
mops_ctx* mops_multi_ctx(mops_multi* mm, int32_t i) { return (mm && i >= 0 && i < mm->n) ? mm->ctx[i] : nullptr; }

int mops_multi_set_mesh(mops_multi* mm, int32_t n_cells, int32_t n_vertices, int32_t max_edges, const double* cell_xyz, const double* vertex_xyz,
                        const int32_t* vertices_on_cell, const int32_t* cells_on_cell, const int32_t* cells_on_vertex, const int32_t* n_edges_on_cell)
{
    if (!mm) return MOPS_E_INVALID;
    return for_devices(mm, [&](int g) {
        return mops_set_mesh(mm->ctx[g], n_cells, n_vertices, max_edges, cell_xyz, vertex_xyz, vertices_on_cell, cells_on_cell, cells_on_vertex,
                             n_edges_on_cell);
    });
}

int mops_multi_set_snapshot(mops_multi* mm, int32_t slot, int32_t n_levels, const double* zonal, const double* meridional,
                            const double* layer_thickness, const double* bottom_depth, const double* vert_vel_top, int32_t n_attr,
                            const double* const* attrs, int32_t n_attr_total, int32_t async)
{
    if (!mm) return MOPS_E_INVALID;
    // every device pulls the snapshot over its own PCIe link and prepares it on its own side stream, in parallel
    return for_devices(mm, [&](int g) {
        return async ? mops_set_snapshot_async(mm->ctx[g], slot, n_levels, zonal, meridional, layer_thickness, bottom_depth, vert_vel_top, n_attr,
                                               attrs, n_attr_total)
                     : mops_set_snapshot(mm->ctx[g], slot, n_levels, zonal, meridional, layer_thickness, bottom_depth, vert_vel_top, n_attr, attrs,
                                         n_attr_total);
    });
}

int mops_multi_snapshot_wait(mops_multi* mm, int32_t slot)
{
    if (!mm) return MOPS_E_INVALID;
    return for_devices(mm, [&](int g) { return mops_snapshot_wait(mm->ctx[g], slot); });
}

int mops_multi_streamline(mops_multi* mm, const mops_traj_cfg* cfg, int32_t slot, const mops_traj_io* io, mops_traj_stats* stats)
{
    return multi_traj(mm, cfg, slot, slot, io, stats, false);
}

int mops_multi_pathline(mops_multi* mm, const mops_traj_cfg* cfg, int32_t front_slot, int32_t back_slot, const mops_traj_io* io,
                        mops_traj_stats* stats)
{
    return multi_traj(mm, cfg, front_slot, back_slot, io, stats, true);
}

} // extern "C"
