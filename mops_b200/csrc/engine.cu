// engine.cu -- host side of the C ABI declared in include/mops_b200.h: resident mesh and
// snapshot management, stream / event plumbing, kernel launches.  No CPU fallback: every
// compute entry point launches sm_100a kernels from kernels.cuh or fails.
#include "../../include/mops_b200.h"
#include "kernels.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <string>
#include <vector>

using namespace mops;

namespace {

struct Snapshot {
    bool valid = false;
    int L = 0;
    int n_attr = 0, n_attr_total = 0;
    double* ztop = nullptr;
    double4* velw = nullptr;
    double* attr[MOPS_MAX_ATTRS] = {nullptr, nullptr};
    unsigned char* mono = nullptr;
    bool has_w = false;   // some velw record has a w component other than +0.0 (vertVelocityTop given and not all zero)
    bool w_known = true;  // false: async upload in flight, the device flag has not been read back yet
    bool w_is_z = false;  // uploaded without vertVelocityTop: the w slot of every velw record holds zTop (SnapView::w_is_z)
    int nonmono = 0;
    size_t bytes = 0;
    cudaEvent_t ready = nullptr;    // recorded on the side stream after preprocessing
    cudaEvent_t last_use = nullptr; // recorded on the main stream after the last kernel reading the slot
    bool pending = false;
    bool used = false;
};

struct Buf { // grow-only device scratch
    void* p = nullptr;
    size_t cap = 0;
};

} // namespace

struct mops_ctx {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr, side = nullptr, own_stream = nullptr;
    cudaEvent_t marks[8] = {};
    cudaStreamAttrValue l2_attr{};
    bool l2_attr_valid = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev_kend = nullptr, ev_end = nullptr;
    std::string err;
    std::atomic<long long> launches{0}; // kernels launched (a snapshot upload may run on a second host thread, see mops_set_snapshot_async)

    // mesh
    bool has_mesh = false;
    int nC = 0, nV = 0, maxEdges = 0, M = 0, F = 0;
    double radius = 0.0;
    void* rec = nullptr;
    double4* c4 = nullptr;
    double4* trig = nullptr;      // caller cell order
    VertRec* vert = nullptr;      // internal vertex order
    int* vcell_ext = nullptr;     // [nV][3] caller cell ids of each internal vertex
    int *c_int2ext = nullptr, *c_ext2int = nullptr, *v_int2ext = nullptr, *v_ext2int = nullptr;
    int* cube = nullptr;
    size_t mesh_bytes = 0;

    Snapshot snap[MOPS_MAX_SNAPSHOT_SLOTS];
    // staging for snapshot upload (caller cell order), reused by every snapshot on the side stream
    Buf st_zonal, st_merid, st_thick, st_wtop, st_bottom, st_ztopc, st_attr, st_vmono;
    int* d_nonmono = nullptr;
    int* d_anyw = nullptr;  // [slot] set by k_vertex_fields when a prepared vertical velocity is not +0.0
    // particle scratch (host-memory mode) + sort scratch
    Buf p_xyz, p_cell0, p_cell_int; // mops_locate (HOST) staging, internal start cells
    Buf s_keys, s_vals, s_keys2, s_vals2, s_tmp;
    Buf s_state;            // AdvState[n]: parked loop state of the compacting multi-launch form
    int* d_nsel = nullptr;  // [2] number of live particles after a compaction (ping-pong: read by the next launch / compaction)
    int segment_steps = 40; // steps per launch of the compacting form; MOPS_SEGMENT_STEPS overrides (0 = one launch per call)
    bool segment_forced = false; // MOPS_SEGMENT_STEPS was given: no adaptation
    double stop_rate = -1.0;     // particles stopped per started step in the last call whose counters were read (< 0: unknown)
    Buf r_img0, r_img1, r_cells;
    unsigned long long* counters = nullptr; // [8]: particle-steps, alive at end, NaN pixels, near-edge particles, above-surface stops
    // HOST-mode trajectory calls: two sets of device staging + events, used alternately, so that the H2D of call k+1
    // (copy_in stream) and the D2H of call k (copy_out stream) overlap the kernels of the other call on the main stream
    struct HostSet {
        Buf xyz, depth, cell0, out_pos, out_vel, out_attr, log, status, steps, fcell, edge;
        unsigned long long* d_counters = nullptr; // [4] device
        unsigned long long* h_counters = nullptr; // [4] pinned host
        cudaEvent_t begin = nullptr, in_done = nullptr, loc0 = nullptr, loc1 = nullptr, k0 = nullptr, k1 = nullptr, ends_done = nullptr,
                    out_done = nullptr;
        bool busy = false;       // a submitted call has not been waited for yet
        long long ticket = 0;    // its ticket
        long long launches = 0;  // kernels it launched
        long long n = 0;         // its particle count
    } hset[2];
    long long host_seq = 0;      // tickets issued so far
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // kd-tree over the cell centres, only for meshes with removed cells (see kd_nearest)
    double4* kd_pts = nullptr;
    unsigned char* kd_dim = nullptr;
    int kd_n = 0;
};

namespace {

int fail(mops_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(ctx, MOPS_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

int ensure(mops_ctx* ctx, Buf& b, size_t bytes)
{
    if (bytes <= b.cap) return MOPS_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(ctx, MOPS_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    b.cap = bytes;
    return MOPS_OK;
}

inline int blocks_for(long long n, int bs) { return (int)((n + bs - 1) / bs); }

// 63-bit Morton key of a point in the unit cube
inline uint64_t spread21(uint64_t x)
{
    x &= 0x1fffffULL;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

// internal numbering = order along a 63-bit Morton curve.  Keys on the host (one pass), the sort on the device: a stable
// LSD radix sort of (key, index) pairs gives the same permutation as std::stable_sort by key (which this replaced: it took
// seconds on the 2.6 M cells + 5.2 M vertices of the level-9 mesh).
int morton_order(mops_ctx* ctx, const double* xyz, int n, std::vector<int>& int2ext)
{
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], xyz[3 * (size_t)i + a]);
            hi[a] = std::max(hi[a], xyz[3 * (size_t)i + a]);
        }
    double ext = 1e-300;
    for (int a = 0; a < 3; ++a) ext = std::max(ext, hi[a] - lo[a]);
    std::vector<uint64_t> key(n);
    auto fill = [&](int b, int e) {
        for (int i = b; i < e; ++i) {
            uint64_t q[3];
            for (int a = 0; a < 3; ++a) {
                double t = (xyz[3 * (size_t)i + a] - lo[a]) / ext;
                t = std::min(std::max(t, 0.0), 1.0);
                q[a] = (uint64_t)(t * 2097151.0);
            }
            key[i] = spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2);
        }
    };
    {
        const int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        std::vector<std::thread> pool;
        const int chunk = (n + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) {
            const int b = t * chunk, e = std::min(n, b + chunk);
            if (b < e) pool.emplace_back(fill, b, e);
        }
        for (auto& th : pool) th.join();
    }
    int2ext.resize(n);
    uint64_t *d_k = nullptr, *d_k2 = nullptr;
    int *d_v = nullptr, *d_v2 = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaStream_t st = ctx->stream;
    auto cleanup = [&] { cudaFree(d_k); cudaFree(d_k2); cudaFree(d_v); cudaFree(d_v2); cudaFree(d_tmp); };
    if (cudaMalloc(&d_k, (size_t)n * 8) != cudaSuccess || cudaMalloc(&d_k2, (size_t)n * 8) != cudaSuccess ||
        cudaMalloc(&d_v, (size_t)n * 4) != cudaSuccess || cudaMalloc(&d_v2, (size_t)n * 4) != cudaSuccess) {
        cleanup();
        return fail(ctx, MOPS_E_NOMEM, "mesh ordering: out of device memory");
    }
    cudaMemcpyAsync(d_k, key.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st);
    k_iota<<<blocks_for(n, 256), 256, 0, st>>>(d_v, n);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_k, d_k2, d_v, d_v2, n, 0, 63, st);
    if (cudaMalloc(&d_tmp, tmp_bytes) != cudaSuccess) {
        cleanup();
        return fail(ctx, MOPS_E_NOMEM, "mesh ordering: out of device memory");
    }
    cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_k, d_k2, d_v, d_v2, n, 0, 63, st);
    cudaMemcpyAsync(int2ext.data(), d_v2, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess) return fail(ctx, MOPS_E_CUDA, "mesh ordering: %s", cudaGetErrorString(e));
    return MOPS_OK;
}

template <int M>
int upload_records(mops_ctx* ctx, const double* vertex_xyz, const int32_t* voc, const int32_t* coc, const int32_t* nedges,
                   const std::vector<int>& c_int2ext, const std::vector<int>& c_ext2int, const std::vector<int>& v_ext2int)
{
    const int nC = ctx->nC, nV = ctx->nV, E = ctx->maxEdges;
    std::vector<CellRec<M>> h((size_t)nC);
    for (int ci = 0; ci < nC; ++ci) {
        const int ce = c_int2ext[ci];
        CellRec<M>& r = h[ci];
        std::memset(&r, 0, sizeof(r));
        int nv = nedges[ce];
        if (nv < 0 || nv > M || nv > E) nv = 0; // unusable cell: every evaluation in it fails (VK:748-751)
        r.nv = nv;
        for (int k = 0; k < M; ++k) {
            r.vid[k] = -1;
            r.nbr[k] = -1;
        }
        for (int k = 0; k < nv; ++k) {
            const int ve = voc[(size_t)ce * E + k] - 1;
            if (ve < 0 || ve >= nV) return fail(ctx, MOPS_E_INVALID, "verticesOnCell[%d][%d] = %d out of range", ce, k, ve + 1);
            r.vid[k] = v_ext2int[ve];
            r.vx[k] = vertex_xyz[3 * (size_t)ve];
            r.vy[k] = vertex_xyz[3 * (size_t)ve + 1];
            r.vz[k] = vertex_xyz[3 * (size_t)ve + 2];
            const int ne = coc[(size_t)ce * E + k] - 1; // 0 pad -> -1 -> skipped, as VK:909-912
            r.nbr[k] = (ne >= 0 && ne < nC) ? c_ext2int[ne] : -1;
        }
    }
    CK(cudaMalloc(&ctx->rec, sizeof(CellRec<M>) * (size_t)nC));
    ctx->mesh_bytes += sizeof(CellRec<M>) * (size_t)nC;
    CK(cudaMemcpyAsync(ctx->rec, h.data(), sizeof(CellRec<M>) * (size_t)nC, cudaMemcpyHostToDevice, ctx->stream));
    k_build_records<M><<<blocks_for(nC, 128), 128, 0, ctx->stream>>>(reinterpret_cast<CellRec<M>*>(ctx->rec), nC);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream)); // h goes out of scope
    return MOPS_OK;
}

template <int M>
int build_cube(mops_ctx* ctx)
{
    // finest resolution: about one bucket per two cells
    int F = 1;
    while (6LL * (2 * F) * (2 * F) <= (long long)ctx->nC / 2 && F < 2048) F *= 2;
    size_t total = 0;
    for (int f = 1; f <= F; f *= 2) total += 6ULL * f * f;
    int* all = nullptr;
    CK(cudaMalloc(&all, total * sizeof(int)));
    ctx->mesh_bytes += total * sizeof(int);
    int* parent = nullptr;
    int* cur = all;
    for (int f = 1; f <= F; f *= 2) {
        k_cube_level<M><<<blocks_for(6LL * f * f, 128), 128, 0, ctx->stream>>>(
            reinterpret_cast<const CellRec<M>*>(ctx->rec), ctx->c4, parent, cur, f, ctx->radius);
        ctx->launches++;
        parent = cur;
        cur += 6ULL * f * f;
    }
    CK(cudaGetLastError());
    ctx->cube = parent; // finest level (coarser levels stay allocated in the same block)
    ctx->F = F;
    return MOPS_OK;
}

void free_mesh(mops_ctx* c)
{
    // the cube table is one allocation whose finest level is the last slice
    if (c->cube) {
        size_t before = 0;
        for (int f = 1; f < c->F; f *= 2) before += 6ULL * f * f;
        cudaFree(c->cube - before);
    }
    cudaFree(c->rec); cudaFree(c->c4); cudaFree(c->trig); cudaFree(c->vert); cudaFree(c->vcell_ext);
    cudaFree(c->c_int2ext); cudaFree(c->c_ext2int); cudaFree(c->v_int2ext); cudaFree(c->v_ext2int);
    cudaFree(c->kd_pts); cudaFree(c->kd_dim);
    c->kd_pts = nullptr; c->kd_dim = nullptr; c->kd_n = 0;
    c->cube = nullptr; c->rec = nullptr; c->c4 = nullptr; c->trig = nullptr; c->vert = nullptr; c->vcell_ext = nullptr;
    c->c_int2ext = c->c_ext2int = c->v_int2ext = c->v_ext2int = nullptr;
    c->has_mesh = false;
    c->mesh_bytes = 0;
}

void free_snapshot(Snapshot& s)
{
    cudaFree(s.ztop); cudaFree(s.velw); cudaFree(s.attr[0]); cudaFree(s.attr[1]); cudaFree(s.mono);
    s.ztop = nullptr; s.velw = nullptr; s.attr[0] = s.attr[1] = nullptr; s.mono = nullptr;
    s.valid = false; s.bytes = 0; s.L = 0;
}

SnapView view_of(const Snapshot& s)
{
    SnapView v;
    v.ztop = s.ztop; v.velw = s.velw; v.attr0 = s.attr[0]; v.attr1 = s.attr[1]; v.mono = s.mono;
    v.w_is_z = s.w_is_z ? 1 : 0;
    return v;
}

int wait_slot(mops_ctx* ctx, int slot)
{
    Snapshot& s = ctx->snap[slot];
    if (s.pending) {
        CK(cudaStreamWaitEvent(ctx->stream, s.ready, 0));
        s.pending = false;
    }
    return MOPS_OK;
}

// a vertVelocityTop array was given with an asynchronous upload: read back whether it held anything but +0.0 (the upload
// of a slot that a trajectory call is about to use was enqueued a whole interval earlier, so this does not stall in a
// double-buffered chain)
int resolve_w(mops_ctx* ctx, int slot)
{
    Snapshot& s = ctx->snap[slot];
    if (s.w_known) return MOPS_OK;
    CK(cudaEventSynchronize(s.ready));
    int any = 1;
    CK(cudaMemcpy(&any, ctx->d_anyw + slot, sizeof(int), cudaMemcpyDeviceToHost));
    s.has_w = any != 0;
    s.w_known = true;
    return MOPS_OK;
}

int mark_use(mops_ctx* ctx, int slot)
{
    Snapshot& s = ctx->snap[slot];
    CK(cudaEventRecord(s.last_use, ctx->stream));
    s.used = true;
    return MOPS_OK;
}

template <int M>
int launch_cell_mono(mops_ctx* ctx, Snapshot& s, int slot, cudaStream_t st)
{
    k_cell_mono<M><<<blocks_for(ctx->nC, 256), 256, 0, st>>>(reinterpret_cast<const CellRec<M>*>(ctx->rec),
                                                             (const unsigned char*)ctx->st_vmono.p, s.mono, ctx->nC, ctx->d_nonmono + slot);
    ctx->launches++;
    return MOPS_OK;
}

int set_snapshot_impl(mops_ctx* ctx, int slot, int L, const double* zonal, const double* merid, const double* thick,
                      const double* bottom, const double* wtop, int n_attr, const double* const* attrs, int n_attr_total,
                      bool async)
{
    if (!ctx) return MOPS_E_INVALID;
    if (!ctx->has_mesh) return fail(ctx, MOPS_E_STATE, "mops_set_snapshot: no mesh");
    if (slot < 0 || slot >= MOPS_MAX_SNAPSHOT_SLOTS) return fail(ctx, MOPS_E_INVALID, "slot %d out of range", slot);
    if (L < 2 || L > 100) return fail(ctx, MOPS_E_INVALID, "n_levels %d outside [2,100] (reference MAX_VERTICAL_LEVEL_NUM)", L);
    if (!zonal || !merid || !thick || !bottom) return fail(ctx, MOPS_E_INVALID, "null snapshot input");
    if (n_attr < 0 || n_attr > MOPS_MAX_ATTRS) return fail(ctx, MOPS_E_INVALID, "n_attr %d outside [0,%d]", n_attr, MOPS_MAX_ATTRS);
    for (int a = 0; a < n_attr; ++a)
        if (!attrs || !attrs[a]) return fail(ctx, MOPS_E_INVALID, "n_attr = %d but attribute array %d is null", n_attr, a);
    if ((long long)ctx->nV * L >= (1LL << 31)) return fail(ctx, MOPS_E_INVALID, "nVertices*nLevels exceeds 2^31");
    CK(cudaSetDevice(ctx->device));
    Snapshot& s = ctx->snap[slot];
    const size_t nC = (size_t)ctx->nC, nV = (size_t)ctx->nV;
    cudaStream_t st = ctx->side;

    // the slot may still be read by kernels on the main stream
    if (s.used) CK(cudaStreamWaitEvent(st, s.last_use, 0));
    if (s.valid && s.L != L) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(st));
        free_snapshot(s);
    }
    if (!s.ztop) {
        CK(cudaMalloc(&s.ztop, nV * L * sizeof(double)));
        CK(cudaMalloc(&s.velw, nV * L * sizeof(double4)));
        CK(cudaMalloc(&s.mono, nC));
        s.bytes = nV * L * (sizeof(double) + sizeof(double4)) + nC;
    }
    for (int a = 0; a < MOPS_MAX_ATTRS; ++a) {
        if (a < n_attr && !s.attr[a]) {
            CK(cudaMalloc(&s.attr[a], nV * L * sizeof(double)));
            s.bytes += nV * L * sizeof(double);
        }
    }
    s.L = L;
    s.has_w = (wtop != nullptr);
    s.w_known = (wtop == nullptr);
    s.w_is_z = (wtop == nullptr);
    s.n_attr = n_attr;
    s.n_attr_total = n_attr_total;

    int rc;
    if ((rc = ensure(ctx, ctx->st_zonal, nC * L * 8))) return rc;
    if ((rc = ensure(ctx, ctx->st_merid, nC * L * 8))) return rc;
    if ((rc = ensure(ctx, ctx->st_thick, nC * L * 8))) return rc;
    if ((rc = ensure(ctx, ctx->st_ztopc, nC * L * 8))) return rc;
    if ((rc = ensure(ctx, ctx->st_bottom, nC * 8))) return rc;
    if ((rc = ensure(ctx, ctx->st_vmono, nV))) return rc;
    if (wtop && (rc = ensure(ctx, ctx->st_wtop, nC * (L + 1) * 8))) return rc;
    if (n_attr > 0 && (rc = ensure(ctx, ctx->st_attr, nC * L * 8))) return rc;

    // inputs are host pointers (the reference's vectors; pinned memory makes the copy asynchronous) or, for callers that
    // assembled the snapshot on the device (e.g. an all-gather over NVLink of per-rank parts), device pointers
    auto kind_of = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return cudaMemcpyHostToDevice; }
        return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    };
    CK(cudaMemcpyAsync(ctx->st_zonal.p, zonal, nC * L * 8, kind_of(zonal), st));
    CK(cudaMemcpyAsync(ctx->st_merid.p, merid, nC * L * 8, kind_of(merid), st));
    CK(cudaMemcpyAsync(ctx->st_thick.p, thick, nC * L * 8, kind_of(thick), st));
    CK(cudaMemcpyAsync(ctx->st_bottom.p, bottom, nC * 8, kind_of(bottom), st));
    if (wtop) CK(cudaMemcpyAsync(ctx->st_wtop.p, wtop, nC * (L + 1) * 8, kind_of(wtop), st));

    CK(cudaMemsetAsync(ctx->d_anyw + slot, 0, sizeof(int), st));
    k_cell_ztop<<<blocks_for(ctx->nC, CZ_CELLS), 128, (size_t)CZ_CELLS * (L + 1) * sizeof(double), st>>>(
        (const double*)ctx->st_thick.p, (const double*)ctx->st_bottom.p, (double*)ctx->st_ztopc.p, ctx->nC, L);
    k_vertex_fields<<<blocks_for((long long)nV * L, 256), 256, 0, st>>>(
        ctx->vert, ctx->vcell_ext, ctx->trig, (const double*)ctx->st_ztopc.p, (const double*)ctx->st_zonal.p,
        (const double*)ctx->st_merid.p, wtop ? (const double*)ctx->st_wtop.p : nullptr, s.ztop, s.velw, ctx->nV, L, ctx->d_anyw + slot,
        wtop ? 0 : 1);
    ctx->launches += 2;
    for (int a = 0; a < n_attr; ++a) {
        CK(cudaMemcpyAsync(ctx->st_attr.p, attrs[a], nC * L * 8, kind_of(attrs[a]), st));
        k_vertex_scalar<<<blocks_for((long long)nV * L, 256), 256, 0, st>>>(ctx->vert, ctx->vcell_ext, (const double*)ctx->st_attr.p,
                                                                          s.attr[a], ctx->nV, L);
        ctx->launches++;
    }
    CK(cudaMemsetAsync(ctx->d_nonmono + slot, 0, sizeof(int), st));
    k_vertex_mono<<<blocks_for((long long)nV * 32, 256), 256, 0, st>>>(s.ztop, (unsigned char*)ctx->st_vmono.p, ctx->nV, L);
    ctx->launches++;
    switch (ctx->M) {
    case 6: launch_cell_mono<6>(ctx, s, slot, st); break;
    case 8: launch_cell_mono<8>(ctx, s, slot, st); break;
    default: launch_cell_mono<20>(ctx, s, slot, st); break;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(s.ready, st));
    s.valid = true;
    s.pending = true;
    s.used = false;
    if (!async) {
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(&s.nonmono, ctx->d_nonmono + slot, sizeof(int), cudaMemcpyDeviceToHost));
        if (!s.w_known) {
            int any = 1;
            CK(cudaMemcpy(&any, ctx->d_anyw + slot, sizeof(int), cudaMemcpyDeviceToHost));
            s.has_w = any != 0;
            s.w_known = true;
        }
        s.pending = false;
    } else {
        s.nonmono = -1;
    }
    return MOPS_OK;
}

inline KdView kd_view(const mops_ctx* ctx)
{
    KdView v;
    v.pts = ctx->kd_pts; v.dim = ctx->kd_dim; v.n = ctx->kd_n;
    return v;
}

template <int M>
void launch_locate(mops_ctx* ctx, long long n, const double* d_xyz, int* d_cell_int, int* d_cell_ext)
{
    k_locate<M><<<blocks_for(n * 8, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const CellRec<M>*>(ctx->rec), ctx->c4, ctx->cube,
                                                                 ctx->F, ctx->nC, kd_view(ctx), n, d_xyz, d_cell_int, d_cell_ext,
                                                                 ctx->c_int2ext);
    ctx->launches++;
}

void dispatch_locate(mops_ctx* ctx, long long n, const double* d_xyz, int* d_cell_int, int* d_cell_ext)
{
    switch (ctx->M) {
    case 6: launch_locate<6>(ctx, n, d_xyz, d_cell_int, d_cell_ext); break;
    case 8: launch_locate<8>(ctx, n, d_xyz, d_cell_int, d_cell_ext); break;
    default: launch_locate<20>(ctx, n, d_xyz, d_cell_int, d_cell_ext); break;
    }
}

// Instantiations of the advection kernel: EXTRA (walk semantics / near-edge diagnostic) and ATTR (pathline
// scalar attributes carried along) are compile-time switches, so the production pathline without attributes
// does not pay registers for either.  Resident 128-thread blocks per SM (register budget 65536 / (128 * MINB)):
// 3 for the 6- and 8-wide records, 1 for the 20-wide ones -- from measurements on B200 (profiles/README.md:
// 2 blocks 990 ms, 3 blocks 817 ms, 4 blocks 855 ms, 5 blocks 1065 ms, 6 blocks 1298 ms on the same step).
template <int M, bool PATH, bool EXTRA, bool ATTR, bool SEG = false, int NOW = 0, bool FAST = false>
void launch_advect_inst(mops_ctx* ctx, const AdvectParams& P)
{
    const int grid = blocks_for(P.n, MOPS_ADV_BLOCK);
    k_advect<M, PATH, (M == 20 ? 1 : MOPS_ADV_MINB), EXTRA, ATTR, SEG, NOW, FAST><<<grid, MOPS_ADV_BLOCK, 0, ctx->stream>>>(P);
    ctx->launches++;
}

// Production launches always use the SEG instantiation (a single launch is the segment [0, times) with no parked state); the
// EXTRA instantiations (walk semantics / near-edge diagnostic) are single-launch only.  Hexagonal meshes (M == 6) without
// attributes run the FAST instantiations (straight-line RK4 step, fastpath.cuh), in the NOW form that matches how the
// snapshots store their vertical velocity (P.no_w).
template <int M, bool PATH>
void launch_fast(mops_ctx* ctx, const AdvectParams& P)
{
    constexpr bool HEX = (M == 6);
    if (HEX && P.no_w == 2) launch_advect_inst<M, PATH, false, false, true, HEX ? 2 : 0, HEX>(ctx, P);
    else if (HEX && P.no_w == 1) launch_advect_inst<M, PATH, false, false, true, HEX ? 1 : 0, HEX>(ctx, P);
    else launch_advect_inst<M, PATH, false, false, true, 0, HEX>(ctx, P);
}

template <int M>
void launch_advect(mops_ctx* ctx, const AdvectParams& P, bool path)
{
    const bool extra = P.walk || P.diag_edge;
    const bool attr = path && P.attr_count > 0 && P.out_attr;
    if (!path) {
        if (extra) launch_advect_inst<M, false, true, false>(ctx, P);
        else launch_fast<M, false>(ctx, P);
    } else if (attr) {
        if (extra) launch_advect_inst<M, true, true, true>(ctx, P);
        else launch_advect_inst<M, true, false, true, true>(ctx, P);
    } else {
        if (extra) launch_advect_inst<M, true, true, false>(ctx, P);
        else launch_fast<M, true>(ctx, P);
    }
}

void dispatch_advect(mops_ctx* ctx, const AdvectParams& P, bool path)
{
    switch (ctx->M) {
    case 6: launch_advect<6>(ctx, P, path); break;
    case 8: launch_advect<8>(ctx, P, path); break;
    default: launch_advect<20>(ctx, P, path); break;
    }
}

// one particle range on the device: start cells (given or located) -> Morton order -> advection kernel, all on
// ctx->stream.  `S` supplies the sort scratch; P carries the device pointers of the range.
int run_range(mops_ctx* ctx, const mops_traj_cfg* cfg, bool path, long long m, const int* d_cell0_ext, int* d_cell_int, Buf& vals,
              Buf& vals2, Buf& keys2, Buf& tmp, AdvectParams& P, cudaEvent_t ev_loc0, cudaEvent_t ev_loc1, cudaEvent_t ev_k0,
              cudaEvent_t ev_k1)
{
    cudaStream_t st = ctx->stream;
    int rc;
    if (ev_loc0) CK(cudaEventRecord(ev_loc0, st));
    if (d_cell0_ext) {
        k_map_ids<<<blocks_for(m, 256), 256, 0, st>>>(d_cell0_ext, d_cell_int, ctx->c_ext2int, ctx->nC, m);
        ctx->launches++;
    } else {
        dispatch_locate(ctx, m, P.pos, d_cell_int, nullptr);
    }
    if (ev_loc1) CK(cudaEventRecord(ev_loc1, st));
    const int* d_order = nullptr;
    if (cfg->sort_particles && m > 1) { // processing order: particles sorted by (Morton-numbered) start cell
        if ((rc = ensure(ctx, vals, (size_t)m * 4))) return rc;
        if ((rc = ensure(ctx, keys2, (size_t)m * 4))) return rc;
        if ((rc = ensure(ctx, vals2, (size_t)m * 4))) return rc;
        k_iota<<<blocks_for(m, 256), 256, 0, st>>>((int*)vals.p, m);
        ctx->launches++;
        size_t tmp_bytes = 0;
        // keys are cell ids in [-1, nC): -1 (invalid) sorts last as unsigned
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned*)d_cell_int, (unsigned*)keys2.p, (const int*)vals.p,
                                        (int*)vals2.p, (int)m, 0, 32, st);
        if ((rc = ensure(ctx, tmp, tmp_bytes))) return rc;
        CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, (const unsigned*)d_cell_int, (unsigned*)keys2.p, (const int*)vals.p,
                                           (int*)vals2.p, (int)m, 0, 32, st));
        d_order = (const int*)vals2.p;
    }
    P.n = m; P.order = d_order; P.cell0 = d_cell_int;
    P.step_begin = 0; P.step_end = P.times; P.state = nullptr; P.n_live = nullptr;
    if (ev_k0) CK(cudaEventRecord(ev_k0, st));
    // Compaction pays when particles stop: a 40-step segment costs ~1.5 % (parked state, one select pass, the launch tails), a
    // stopped particle's lane costs its share of every remaining step.  The rate at which the previous call's particles
    // stopped predicts this one (chained intervals, repeated calls): below one stop per 4000 started steps (1.5 % of the
    // particles over a 60-step half interval) the call runs as one launch.  Results do not depend on the choice.
    int seg = ctx->segment_steps;
    if (!ctx->segment_forced && ctx->stop_rate >= 0.0 && ctx->stop_rate < 2.5e-4) seg = 0;
    if (seg > 0 && P.times > seg && !(P.walk || P.diag_edge)) {
        // Compacting multi-launch form (default 40 steps per launch, MOPS_SEGMENT_STEPS=<steps>, 0 = off): under the reference's semantics particles
        // stop for good at their first failed stage, and a stopped particle's lane idles until its whole warp is
        // done.  Each launch integrates `seg` steps; particles still alive park their loop state, the processing
        // order is compacted to them (order-preserving, so the Morton locality stays) and the next launch carries
        // only live lanes.  Results are identical to the single launch: same per-particle arithmetic, and every
        // output slot is still written exactly once by the launch in which the particle stops or finishes.
        if ((rc = ensure(ctx, ctx->s_state, (size_t)m * sizeof(AdvState)))) return rc;
        if ((rc = ensure(ctx, vals, (size_t)m * 4))) return rc;
        if ((rc = ensure(ctx, vals2, (size_t)m * 4))) return rc;
        int* cur = (int*)vals2.p; // the sorted order lives here
        int* nxt = (int*)vals.p;
        if (!d_order) { // unsorted call: start from the identity order
            k_iota<<<blocks_for(m, 256), 256, 0, st>>>(cur, m);
            ctx->launches++;
        }
        P.state = (AdvState*)ctx->s_state.p;
        // The live count stays on the device (d_nsel[0/1], written by one compaction and read by the next launch and the
        // next compaction): every launch is sized for all m positions and lanes beyond the live prefix exit at once, so the
        // host never waits for a segment and the whole call is stream-asynchronous (and graph-capturable).
        const int* n_live = nullptr;
        int flip = 0;
        for (int b = 0; b < P.times; b += seg) {
            const int e = std::min(P.times, b + seg);
            P.n = m; P.order = cur; P.step_begin = b; P.step_end = e; P.n_live = n_live;
            dispatch_advect(ctx, P, path);
            CK(cudaGetLastError());
            if (e < P.times) {
                size_t tmp_bytes = 0;
                const AdvAliveFlag op{cur, P.state, n_live};
                cub::TransformInputIterator<bool, AdvAliveFlag, cub::CountingInputIterator<int>> flags(cub::CountingInputIterator<int>(0), op);
                int* n_out = ctx->d_nsel + flip;
                cub::DeviceSelect::Flagged(nullptr, tmp_bytes, (const int*)cur, flags, nxt, n_out, (int)m, st);
                if ((rc = ensure(ctx, tmp, tmp_bytes))) return rc;
                CK(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, (const int*)cur, flags, nxt, n_out, (int)m, st));
                n_live = n_out;
                flip ^= 1;
                std::swap(cur, nxt);
            }
        }
    } else {
        dispatch_advect(ctx, P, path);
        CK(cudaGetLastError());
    }
    if (ev_k1) CK(cudaEventRecord(ev_k1, st));
    return MOPS_OK;
}

// wait for a submitted HOST-mode call: what = 0 -> its end points / depths have landed in io.xyz / io.depth,
// what = 1 -> every output has; stats are filled when the whole call is complete
int host_wait(mops_ctx* ctx, long long ticket, int what, mops_traj_stats* stats)
{
    mops_ctx::HostSet& S = ctx->hset[ticket & 1];
    if (ticket <= 0 || ticket > ctx->host_seq) return fail(ctx, MOPS_E_INVALID, "unknown ticket %lld", ticket);
    if (S.ticket != ticket) { // already completed (and its staging set reused) -- nothing left to wait for
        if (stats) std::memset(stats, 0, sizeof(*stats));
        return MOPS_OK;
    }
    CK(cudaSetDevice(ctx->device));
    if (what == 0) {
        CK(cudaEventSynchronize(S.ends_done));
        return MOPS_OK;
    }
    CK(cudaEventSynchronize(S.out_done));
    S.busy = false;
    if (S.h_counters[0] > 0)
        ctx->stop_rate = (double)((unsigned long long)S.n - std::min<unsigned long long>(S.h_counters[1], (unsigned long long)S.n)) / (double)S.h_counters[0];
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->particle_steps = (int64_t)S.h_counters[0];
        stats->alive_at_end = (int64_t)S.h_counters[1];
        stats->near_edge_particles = (int64_t)S.h_counters[3];
        stats->above_surface_particles = (int64_t)S.h_counters[4];
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, S.k0, S.k1) == cudaSuccess) stats->kernel_ms = ms;
        if (cudaEventElapsedTime(&ms, S.loc0, S.loc1) == cudaSuccess) stats->locate_ms = ms;
        if (cudaEventElapsedTime(&ms, S.begin, S.out_done) == cudaSuccess) stats->total_ms = ms;
        stats->launches = (int32_t)S.launches;
    }
    return MOPS_OK;
}

// ticket != nullptr: HOST-mode submit -- returns once everything is enqueued; host_wait completes it.
int trajectory_impl(mops_ctx* ctx, const mops_traj_cfg* cfg, int front, int back, const mops_traj_io* io, mops_traj_stats* stats,
                    bool path, long long* ticket)
{
    if (!ctx) return MOPS_E_INVALID;
    if (ticket) *ticket = 0;
    if (!cfg || !io) return fail(ctx, MOPS_E_INVALID, "null cfg/io");
    if (!ctx->has_mesh) return fail(ctx, MOPS_E_STATE, "no mesh");
    if (front < 0 || front >= MOPS_MAX_SNAPSHOT_SLOTS || !ctx->snap[front].valid) return fail(ctx, MOPS_E_STATE, "snapshot slot %d not set", front);
    if (path && (back < 0 || back >= MOPS_MAX_SNAPSHOT_SLOTS || !ctx->snap[back].valid)) return fail(ctx, MOPS_E_STATE, "snapshot slot %d not set", back);
    // reference: Error("invalid trajectory settings") + empty result (VK:666-669, 1030-1033)
    if (cfg->delta_t <= 0 || cfg->record_t <= 0 || cfg->duration <= 0) return fail(ctx, MOPS_E_INVALID, "invalid trajectory settings");
    if (cfg->delta_t > 0x7fffffffLL || cfg->record_t > 0x7fffffffLL || cfg->duration > 0x7fffffffLL)
        return fail(ctx, MOPS_E_INVALID, "trajectory settings exceed the reference's int range");
    const long long n = io->n;
    if (n < 0 || n > 0x7fffffffLL) return fail(ctx, MOPS_E_INVALID, "particle count out of range");
    const int each = (int)(cfg->duration / cfg->record_t);
    const int times = (int)(cfg->duration / cfg->delta_t);
    if (each <= 0 || times <= 0) return fail(ctx, MOPS_E_INVALID, "invalid integration steps"); // VK:709-712
    const bool host = (cfg->mem == MOPS_MEM_HOST);
    if (ticket && !host) return fail(ctx, MOPS_E_INVALID, "submit/wait is the HOST-memory form; DEVICE-memory calls are asynchronous on the context's stream already");
    if (n == 0) {
        if (stats) std::memset(stats, 0, sizeof(*stats));
        return MOPS_OK;
    }
    if (!io->xyz || !io->depth || !io->out_pos || !io->out_vel) return fail(ctx, MOPS_E_INVALID, "null particle buffers");
    Snapshot& F = ctx->snap[front];
    Snapshot& B = ctx->snap[path ? back : front];
    if (path && F.L != B.L) return fail(ctx, MOPS_E_INVALID, "front/back level counts differ");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = wait_slot(ctx, front))) return rc;
    if (path && (rc = wait_slot(ctx, back))) return rc;
    if ((rc = resolve_w(ctx, front))) return rc;
    if (path && (rc = resolve_w(ctx, back))) return rc;

    const long long launches0 = ctx->launches;
    cudaStream_t st = ctx->stream;

    // pathline attributes: only when the front snapshot holds more than one (VK:1093-1104)
    int attr_count = 0;
    if (path && F.n_attr_total > 1) attr_count = std::min(std::min(F.n_attr, B.n_attr), MOPS_MAX_ATTRS);
    const bool want_attr = path && attr_count > 0 && io->out_attr;

    AdvectParams P;
    std::memset(&P, 0, sizeof(P));
    P.rec = ctx->rec; P.c4 = ctx->c4; P.c_int2ext = ctx->c_int2ext; P.nC = ctx->nC; P.L = F.L;
    P.sv[0] = view_of(F); P.sv[1] = view_of(B);
    P.attr_count = attr_count;
    P.no_w = (!F.has_w && !B.has_w) ? ((F.w_is_z && B.w_is_z) ? 2 : 1) : 0;
    P.use_euler = (cfg->method == MOPS_METHOD_EULER) ? 1 : 0;
    P.delta_t = (cfg->direction == MOPS_DIR_FORWARD ? 1 : -1) * (int)cfg->delta_t;
    P.times = times; P.each = each; P.record_t = (int)cfg->record_t;
    P.record_interval = (int)(cfg->record_t / cfg->delta_t);
    P.duration = (double)cfg->duration;
    P.walk = (cfg->semantics == MOPS_SEM_WALK) ? 1 : 0;

    const size_t out_bytes = (size_t)n * each * 3 * sizeof(double);
    if ((rc = ensure(ctx, ctx->p_cell_int, (size_t)n * 4))) return rc;

    if (!host) {
        // DEVICE memory: everything on the context's stream, asynchronous unless the caller asks for stats
        CK(cudaEventRecord(ctx->ev0, st));
        if (io->out_cell_log) CK(cudaMemsetAsync(io->out_cell_log, 0xff, (size_t)n * times * 4, st));
        CK(cudaMemsetAsync(ctx->counters, 0, 8 * sizeof(unsigned long long), st));
        P.counters = ctx->counters;
        P.pos = io->xyz; P.depth = io->depth;
        P.out_pos = io->out_pos; P.out_vel = io->out_vel; P.out_attr = want_attr ? io->out_attr : nullptr;
        P.cell_log = io->out_cell_log; P.status = io->out_status; P.steps = io->out_steps; P.fcell = io->out_cell;
        P.min_edge = io->out_min_edge;
        P.diag_edge = (cfg->count_near_edge || io->out_min_edge) ? 1 : 0;
        if ((rc = run_range(ctx, cfg, path, n, io->cell0, (int*)ctx->p_cell_int.p, ctx->s_vals, ctx->s_vals2, ctx->s_keys2, ctx->s_tmp, P,
                            ctx->ev2, ctx->ev3, ctx->ev1, ctx->ev_kend)))
            return rc;
        if ((rc = mark_use(ctx, front))) return rc;
        if (path && back != front && (rc = mark_use(ctx, back))) return rc;
        if (stats) {
            unsigned long long h_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            CK(cudaMemcpyAsync(h_counters, ctx->counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ctx->ev_end, st));
            CK(cudaStreamSynchronize(st));
            std::memset(stats, 0, sizeof(*stats));
            stats->particle_steps = (int64_t)h_counters[0];
            stats->alive_at_end = (int64_t)h_counters[1];
            stats->near_edge_particles = (int64_t)h_counters[3];
            stats->above_surface_particles = (int64_t)h_counters[4];
            if (h_counters[0] > 0) ctx->stop_rate = (double)((unsigned long long)n - std::min<unsigned long long>(h_counters[1], (unsigned long long)n)) / (double)h_counters[0];
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev_kend) == cudaSuccess) stats->kernel_ms = ms;
            if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) stats->locate_ms = ms;
            if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev_end) == cudaSuccess) stats->total_ms = ms;
            stats->launches = (int32_t)(ctx->launches - launches0);
        }
        return MOPS_OK;
    }

    // HOST memory: inputs go up on the copy_in stream into one of two staging sets, the kernels run on the main stream,
    // end points first and then the recorded trajectories come back on the copy_out stream.  A second call submitted
    // before the first is waited for uses the other set: its H2D and the first call's D2H overlap the kernels.
    const long long tk = ++ctx->host_seq;
    mops_ctx::HostSet& S = ctx->hset[tk & 1];
    if (S.busy) { // the set's previous call (ticket tk - 2) was never waited for: finish it before its staging is reused
        CK(cudaEventSynchronize(S.out_done));
        S.busy = false;
    }
    if ((rc = ensure(ctx, S.xyz, (size_t)n * 24))) return rc;
    if ((rc = ensure(ctx, S.depth, (size_t)n * 4))) return rc;
    if ((rc = ensure(ctx, S.out_pos, out_bytes))) return rc;
    if ((rc = ensure(ctx, S.out_vel, out_bytes))) return rc;
    if (io->cell0 && (rc = ensure(ctx, S.cell0, (size_t)n * 4))) return rc;
    if (want_attr && (rc = ensure(ctx, S.out_attr, out_bytes))) return rc;
    if (io->out_cell_log && (rc = ensure(ctx, S.log, (size_t)n * times * 4))) return rc;
    if (io->out_status && (rc = ensure(ctx, S.status, (size_t)n * 4))) return rc;
    if (io->out_steps && (rc = ensure(ctx, S.steps, (size_t)n * 4))) return rc;
    if (io->out_cell && (rc = ensure(ctx, S.fcell, (size_t)n * 4))) return rc;
    if (io->out_min_edge && (rc = ensure(ctx, S.edge, (size_t)n * 8))) return rc;

    cudaStream_t si = ctx->copy_in, so = ctx->copy_out;
    CK(cudaEventRecord(S.begin, si));
    CK(cudaMemcpyAsync(S.xyz.p, io->xyz, (size_t)n * 24, cudaMemcpyHostToDevice, si));
    CK(cudaMemcpyAsync(S.depth.p, io->depth, (size_t)n * 4, cudaMemcpyHostToDevice, si));
    if (io->cell0) CK(cudaMemcpyAsync(S.cell0.p, io->cell0, (size_t)n * 4, cudaMemcpyHostToDevice, si));
    CK(cudaEventRecord(S.in_done, si));

    CK(cudaStreamWaitEvent(st, S.in_done, 0));
    if (io->out_cell_log) CK(cudaMemsetAsync(S.log.p, 0xff, (size_t)n * times * 4, st));
    CK(cudaMemsetAsync(S.d_counters, 0, 8 * sizeof(unsigned long long), st));
    P.counters = S.d_counters;
    P.pos = (double*)S.xyz.p; P.depth = (float*)S.depth.p;
    P.out_pos = (double*)S.out_pos.p; P.out_vel = (double*)S.out_vel.p; P.out_attr = want_attr ? (double*)S.out_attr.p : nullptr;
    P.cell_log = io->out_cell_log ? (int*)S.log.p : nullptr;
    P.status = io->out_status ? (int*)S.status.p : nullptr;
    P.steps = io->out_steps ? (int*)S.steps.p : nullptr;
    P.fcell = io->out_cell ? (int*)S.fcell.p : nullptr;
    P.min_edge = io->out_min_edge ? (double*)S.edge.p : nullptr;
    P.diag_edge = (cfg->count_near_edge || P.min_edge) ? 1 : 0;
    if ((rc = run_range(ctx, cfg, path, n, io->cell0 ? (const int*)S.cell0.p : nullptr, (int*)ctx->p_cell_int.p, ctx->s_vals, ctx->s_vals2,
                        ctx->s_keys2, ctx->s_tmp, P, S.loc0, S.loc1, S.k0, S.k1)))
        return rc;
    if ((rc = mark_use(ctx, front))) return rc;
    if (path && back != front && (rc = mark_use(ctx, back))) return rc;

    CK(cudaStreamWaitEvent(so, S.k1, 0));
    CK(cudaMemcpyAsync(io->xyz, S.xyz.p, (size_t)n * 24, cudaMemcpyDeviceToHost, so));
    CK(cudaMemcpyAsync(io->depth, S.depth.p, (size_t)n * 4, cudaMemcpyDeviceToHost, so));
    CK(cudaEventRecord(S.ends_done, so));
    CK(cudaMemcpyAsync(io->out_pos, S.out_pos.p, out_bytes, cudaMemcpyDeviceToHost, so));
    CK(cudaMemcpyAsync(io->out_vel, S.out_vel.p, out_bytes, cudaMemcpyDeviceToHost, so));
    if (want_attr) CK(cudaMemcpyAsync(io->out_attr, S.out_attr.p, out_bytes, cudaMemcpyDeviceToHost, so));
    if (io->out_cell_log) CK(cudaMemcpyAsync(io->out_cell_log, S.log.p, (size_t)n * times * 4, cudaMemcpyDeviceToHost, so));
    if (io->out_status) CK(cudaMemcpyAsync(io->out_status, S.status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, so));
    if (io->out_steps) CK(cudaMemcpyAsync(io->out_steps, S.steps.p, (size_t)n * 4, cudaMemcpyDeviceToHost, so));
    if (io->out_cell) CK(cudaMemcpyAsync(io->out_cell, S.fcell.p, (size_t)n * 4, cudaMemcpyDeviceToHost, so));
    if (io->out_min_edge) CK(cudaMemcpyAsync(io->out_min_edge, S.edge.p, (size_t)n * 8, cudaMemcpyDeviceToHost, so));
    CK(cudaMemcpyAsync(S.h_counters, S.d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, so));
    CK(cudaEventRecord(S.out_done, so));
    S.busy = true;
    S.ticket = tk;
    S.launches = ctx->launches.load() - launches0;
    S.n = n;
    if (ticket) {
        *ticket = tk;
        return MOPS_OK;
    }
    return host_wait(ctx, tk, 1, stats); // the plain HOST-mode call completes before it returns
}

template <int M>
void launch_remap(mops_ctx* ctx, const RemapParams& P)
{
    k_remap<M><<<blocks_for((long long)P.width * P.height, 128), 128, 0, ctx->stream>>>(P);
    ctx->launches++;
}

template <int M>
void launch_view(mops_ctx* ctx, const ViewParams& P, int mode)
{
    const int grid = blocks_for((long long)P.width * P.height, 128);
    if (mode == 0) k_view<M, 0><<<grid, 128, 0, ctx->stream>>>(P);
    else k_view<M, 1><<<grid, 128, 0, ctx->stream>>>(P);
    ctx->launches++;
}

int view_impl(mops_ctx* ctx, const mops_view_cfg* cfg, int slot, double* img, int* pixel_cell, mops_remap_stats* stats, int mode)
{
    if (!ctx) return MOPS_E_INVALID;
    if (!cfg || !img) return fail(ctx, MOPS_E_INVALID, "null cfg/img");
    if (!ctx->has_mesh) return fail(ctx, MOPS_E_STATE, "no mesh");
    if (slot < 0 || slot >= MOPS_MAX_SNAPSHOT_SLOTS || !ctx->snap[slot].valid) return fail(ctx, MOPS_E_STATE, "snapshot slot %d not set", slot);
    if (cfg->width <= 0 || cfg->height <= 0) return fail(ctx, MOPS_E_INVALID, "invalid image size"); // VK:160-163, 483-486
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = wait_slot(ctx, slot))) return rc;
    Snapshot& S = ctx->snap[slot];
    const size_t npx = (size_t)cfg->width * cfg->height;
    const bool host = (cfg->mem == MOPS_MEM_HOST);
    const long long launches0 = ctx->launches;
    cudaStream_t st = ctx->stream;
    double* d0;
    int* dc = nullptr;
    CK(cudaEventRecord(ctx->ev0, st));
    if (host) {
        if ((rc = ensure(ctx, ctx->r_img0, npx * 32))) return rc;
        d0 = (double*)ctx->r_img0.p;
        if (pixel_cell) { if ((rc = ensure(ctx, ctx->r_cells, npx * 4))) return rc; dc = (int*)ctx->r_cells.p; }
    } else {
        d0 = img; dc = pixel_cell;
    }
    CK(cudaMemsetAsync(ctx->counters, 0, 8 * sizeof(unsigned long long), st));
    ViewParams P;
    std::memset(&P, 0, sizeof(P));
    P.rec = ctx->rec; P.c4 = ctx->c4; P.cube = ctx->cube; P.c_int2ext = ctx->c_int2ext; P.F = ctx->F; P.nC = ctx->nC; P.L = S.L;
    P.kd = kd_view(ctx);
    P.s = view_of(S);
    P.width = cfg->width; P.height = cfg->height;
    P.minLat = cfg->lat_min; P.maxLat = cfg->lat_max; P.minLon = cfg->lon_min; P.maxLon = cfg->lon_max;
    P.fixed_layer = std::min(std::max(cfg->fixed_layer, 0), S.L - 1); // ClampLayer, VK:14-26
    P.fixed_lat = cfg->fixed_latitude; P.minDepth = cfg->depth_min; P.maxDepth = cfg->depth_max;
    P.img = d0; P.pixel_cell = dc; P.nan_count = ctx->counters + 2;
    CK(cudaEventRecord(ctx->ev1, st));
    switch (ctx->M) {
    case 6: launch_view<6>(ctx, P, mode); break;
    case 8: launch_view<8>(ctx, P, mode); break;
    default: launch_view<20>(ctx, P, mode); break;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev2, st));
    if ((rc = mark_use(ctx, slot))) return rc;
    unsigned long long h_counters[4] = {0, 0, 0, 0};
    if (host) {
        CK(cudaMemcpyAsync(img, d0, npx * 32, cudaMemcpyDeviceToHost, st));
        if (dc) CK(cudaMemcpyAsync(pixel_cell, dc, npx * 4, cudaMemcpyDeviceToHost, st));
    }
    if (host || stats) {
        CK(cudaMemcpyAsync(h_counters, ctx->counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev3, st));
        CK(cudaStreamSynchronize(st));
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev2) == cudaSuccess) stats->kernel_ms = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev3) == cudaSuccess) stats->total_ms = ms;
        stats->nan_pixels = (int64_t)h_counters[2];
        stats->launches = (int32_t)(ctx->launches - launches0);
        stats->n_images = 1;
    }
    return MOPS_OK;
}

} // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int mops_abi_version(void) { return MOPS_B200_ABI_VERSION; }

int mops_create(mops_ctx** out, int device_ordinal)
{
    if (!out) return MOPS_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return MOPS_E_NODEVICE;
    if (device_ordinal < 0 || device_ordinal >= ndev) return MOPS_E_NODEVICE;
    mops_ctx* ctx = new mops_ctx();
    ctx->device = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device_ordinal) != cudaSuccess) {
        delete ctx;
        return MOPS_E_NODEVICE;
    }
    // the kernels are built for sm_100a only: refuse anything else loudly
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, (const void*)k_iota) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return MOPS_E_NODEVICE;
    }
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess &&
              cudaEventCreate(&ctx->ev2) == cudaSuccess && cudaEventCreate(&ctx->ev3) == cudaSuccess &&
              cudaEventCreate(&ctx->ev_kend) == cudaSuccess && cudaEventCreate(&ctx->ev_end) == cudaSuccess &&
              cudaMalloc(&ctx->counters, 8 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&ctx->d_nonmono, MOPS_MAX_SNAPSHOT_SLOTS * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&ctx->d_anyw, MOPS_MAX_SNAPSHOT_SLOTS * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&ctx->d_nsel, 2 * sizeof(int)) == cudaSuccess;
    if (const char* e = getenv("MOPS_SEGMENT_STEPS")) { ctx->segment_steps = std::max(0, atoi(e)); ctx->segment_forced = true; }
    ctx->stream = ctx->own_stream;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 2; ++i) {
        auto& S = ctx->hset[i];
        ok = cudaEventCreate(&S.begin) == cudaSuccess && cudaEventCreateWithFlags(&S.in_done, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreate(&S.loc0) == cudaSuccess && cudaEventCreate(&S.loc1) == cudaSuccess && cudaEventCreate(&S.k0) == cudaSuccess &&
             cudaEventCreate(&S.k1) == cudaSuccess && cudaEventCreateWithFlags(&S.ends_done, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreate(&S.out_done) == cudaSuccess && cudaMalloc(&S.d_counters, 8 * sizeof(unsigned long long)) == cudaSuccess &&
             cudaHostAlloc(&S.h_counters, 8 * sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
    }
    for (int i = 0; ok && i < 8; ++i) ok = cudaEventCreate(&ctx->marks[i]) == cudaSuccess;
    for (int i = 0; ok && i < MOPS_MAX_SNAPSHOT_SLOTS; ++i) {
        ok = cudaEventCreate(&ctx->snap[i].ready) == cudaSuccess && cudaEventCreate(&ctx->snap[i].last_use) == cudaSuccess;
    }
    if (!ok) {
        delete ctx;
        return MOPS_E_CUDA;
    }
    *out = ctx;
    return MOPS_OK;
}

void mops_destroy(mops_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_mesh(ctx);
    for (auto& s : ctx->snap) {
        free_snapshot(s);
        if (s.ready) cudaEventDestroy(s.ready);
        if (s.last_use) cudaEventDestroy(s.last_use);
    }
    Buf* bufs[] = {&ctx->st_zonal, &ctx->st_merid, &ctx->st_thick, &ctx->st_wtop, &ctx->st_bottom, &ctx->st_ztopc, &ctx->st_attr,
                   &ctx->st_vmono, &ctx->p_xyz, &ctx->p_cell0, &ctx->p_cell_int, &ctx->s_keys, &ctx->s_vals,
                   &ctx->s_keys2, &ctx->s_vals2, &ctx->s_tmp, &ctx->s_state, &ctx->r_img0, &ctx->r_img1, &ctx->r_cells};
    for (Buf* b : bufs) cudaFree(b->p);
    for (auto& S : ctx->hset) {
        Buf* pb[] = {&S.xyz, &S.depth, &S.cell0, &S.out_pos, &S.out_vel, &S.out_attr, &S.log, &S.status, &S.steps, &S.fcell, &S.edge};
        for (Buf* b : pb) cudaFree(b->p);
        cudaFree(S.d_counters);
        if (S.h_counters) cudaFreeHost(S.h_counters);
        cudaEvent_t evs[] = {S.begin, S.in_done, S.loc0, S.loc1, S.k0, S.k1, S.ends_done, S.out_done};
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    }
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    cudaFree(ctx->counters);
    cudaFree(ctx->d_nonmono);
    cudaFree(ctx->d_anyw);
    cudaFree(ctx->d_nsel);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev2); cudaEventDestroy(ctx->ev3);
    cudaEventDestroy(ctx->ev_kend); cudaEventDestroy(ctx->ev_end);
    for (auto& m : ctx->marks) if (m) cudaEventDestroy(m);
    cudaStreamDestroy(ctx->own_stream);
    cudaStreamDestroy(ctx->side);
    delete ctx;
}

const char* mops_last_error(const mops_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mops_host_alloc(void** out, size_t bytes)
{
    if (!out) return MOPS_E_INVALID;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? MOPS_OK : MOPS_E_NOMEM;
}

int mops_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? MOPS_OK : MOPS_E_CUDA; }

int mops_synchronize(mops_ctx* ctx)
{
    if (!ctx) return MOPS_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->side));
    CK(cudaStreamSynchronize(ctx->stream));
    return MOPS_OK;
}

int mops_set_stream(mops_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return MOPS_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    if (ctx->l2_attr_valid) {
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &ctx->l2_attr);
        cudaGetLastError();
    }
    return MOPS_OK;
}

int mops_mark(mops_ctx* ctx, int32_t idx)
{
    if (!ctx || idx < 0 || idx >= 8) return MOPS_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->marks[idx], ctx->stream));
    return MOPS_OK;
}

int mops_elapsed_ms(mops_ctx* ctx, int32_t idx_from, int32_t idx_to, double* ms_out)
{
    if (!ctx || !ms_out || idx_from < 0 || idx_from >= 8 || idx_to < 0 || idx_to >= 8) return MOPS_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->marks[idx_to]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->marks[idx_from], ctx->marks[idx_to]));
    *ms_out = ms;
    return MOPS_OK;
}

// median-split kd-tree over the cell centres (internal numbering) in the implicit layout kd_nearest walks
static void build_kd(const std::vector<double4>& c4, std::vector<double4>& pts, std::vector<unsigned char>& dims)
{
    const int n = (int)c4.size();
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    dims.assign(n, 0);
    std::vector<std::pair<int, int>> todo;
    todo.emplace_back(0, n);
    while (!todo.empty()) {
        const int lo = todo.back().first, hi = todo.back().second;
        todo.pop_back();
        if (hi - lo <= 0) continue;
        const int mid = (lo + hi) >> 1;
        if (hi - lo > 1) {
            double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
            for (int i = lo; i < hi; ++i) {
                const double4& p = c4[idx[i]];
                const double v[3] = {p.x, p.y, p.z};
                for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); }
            }
            int ax = 0;
            if (mx[1] - mn[1] > mx[ax] - mn[ax]) ax = 1;
            if (mx[2] - mn[2] > mx[ax] - mn[ax]) ax = 2;
            std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
                const double va = ax == 0 ? c4[a].x : ax == 1 ? c4[a].y : c4[a].z;
                const double vb = ax == 0 ? c4[b].x : ax == 1 ? c4[b].y : c4[b].z;
                return va < vb;
            });
            dims[mid] = (unsigned char)ax;
            todo.emplace_back(lo, mid);
            todo.emplace_back(mid + 1, hi);
        }
    }
    pts.resize(n);
    for (int i = 0; i < n; ++i) {
        const double4& p = c4[idx[i]];
        long long id = idx[i];
        double w;
        std::memcpy(&w, &id, 8);
        pts[i] = make_double4(p.x, p.y, p.z, w);
    }
}

int mops_set_mesh(mops_ctx* ctx, int32_t n_cells, int32_t n_vertices, int32_t max_edges, const double* cell_xyz,
                  const double* vertex_xyz, const int32_t* vertices_on_cell, const int32_t* cells_on_cell,
                  const int32_t* cells_on_vertex, const int32_t* n_edges_on_cell)
{
    if (!ctx) return MOPS_E_INVALID;
    if (n_cells <= 0 || n_vertices <= 0 || max_edges <= 0) return fail(ctx, MOPS_E_INVALID, "bad mesh sizes");
    if (!cell_xyz || !vertex_xyz || !vertices_on_cell || !cells_on_cell || !cells_on_vertex || !n_edges_on_cell)
        return fail(ctx, MOPS_E_INVALID, "null mesh array");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    free_mesh(ctx);
    for (auto& s : ctx->snap) free_snapshot(s);
    ctx->nC = n_cells; ctx->nV = n_vertices; ctx->maxEdges = max_edges;
    const size_t nC = (size_t)n_cells, nV = (size_t)n_vertices;

    int max_nv = 0;
    for (size_t c = 0; c < nC; ++c) max_nv = std::max(max_nv, (int)n_edges_on_cell[c]);
    max_nv = std::min(max_nv, (int)max_edges);
    // reference: more than MAX_VERTEX_NUM = 20 vertices makes every evaluation fail (VK:741-751)
    ctx->M = (max_nv <= 6) ? 6 : (max_nv <= 8) ? 8 : 20;

    // Morton renumbering of cells and vertices (layout only; ids crossing the ABI stay the caller's)
    std::vector<int> c_int2ext, v_int2ext, c_ext2int(nC), v_ext2int(nV);
    {
        int rc0;
        if ((rc0 = morton_order(ctx, cell_xyz, n_cells, c_int2ext))) return rc0;
        if ((rc0 = morton_order(ctx, vertex_xyz, n_vertices, v_int2ext))) return rc0;
    }
    for (size_t i = 0; i < nC; ++i) c_ext2int[c_int2ext[i]] = (int)i;
    for (size_t i = 0; i < nV; ++i) v_ext2int[v_int2ext[i]] = (int)i;

    double rsum = 0.0;
    const size_t rs = std::min<size_t>(nC, 64);
    for (size_t i = 0; i < rs; ++i)
        rsum += std::sqrt(cell_xyz[3 * i] * cell_xyz[3 * i] + cell_xyz[3 * i + 1] * cell_xyz[3 * i + 1] + cell_xyz[3 * i + 2] * cell_xyz[3 * i + 2]);
    ctx->radius = rsum / (double)rs;

    int rc = MOPS_OK;
    switch (ctx->M) {
    case 6: rc = upload_records<6>(ctx, vertex_xyz, vertices_on_cell, cells_on_cell, n_edges_on_cell, c_int2ext, c_ext2int, v_ext2int); break;
    case 8: rc = upload_records<8>(ctx, vertex_xyz, vertices_on_cell, cells_on_cell, n_edges_on_cell, c_int2ext, c_ext2int, v_ext2int); break;
    default: rc = upload_records<20>(ctx, vertex_xyz, vertices_on_cell, cells_on_cell, n_edges_on_cell, c_int2ext, c_ext2int, v_ext2int); break;
    }
    if (rc) { free_mesh(ctx); return rc; }

    // cell centres (internal order, one 32 B sector each), id maps
    std::vector<double4> h_c4(nC);
    for (size_t i = 0; i < nC; ++i) {
        const size_t e = (size_t)c_int2ext[i];
        h_c4[i] = make_double4(cell_xyz[3 * e], cell_xyz[3 * e + 1], cell_xyz[3 * e + 2], 0.0);
    }
    // culled mesh (a neighbour of some cell was removed)?  Then the greedy walk is not exact on its own.
    bool culled = false;
    for (size_t c = 0; c < nC && !culled; ++c) {
        const int nv = std::min((int)n_edges_on_cell[c], (int)max_edges);
        for (int k = 0; k < nv; ++k) {
            const long long e = (long long)cells_on_cell[c * (size_t)max_edges + k] - 1;
            if (e < 0 || e >= (long long)n_cells) { culled = true; break; }
        }
    }
    if (culled) {
        std::vector<double4> kd_pts;
        std::vector<unsigned char> kd_dim;
        build_kd(h_c4, kd_pts, kd_dim);
        CK(cudaMalloc(&ctx->kd_pts, nC * sizeof(double4)));
        CK(cudaMalloc(&ctx->kd_dim, nC));
        CK(cudaMemcpy(ctx->kd_pts, kd_pts.data(), nC * sizeof(double4), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->kd_dim, kd_dim.data(), nC, cudaMemcpyHostToDevice));
        ctx->kd_n = n_cells;
        ctx->mesh_bytes += nC * (sizeof(double4) + 1);
    }
    std::vector<VertRec> h_vert(nV);
    std::vector<int> h_vcell(nV * 3);
    for (size_t i = 0; i < nV; ++i) {
        const size_t ve = (size_t)v_int2ext[i];
        VertRec r;
        std::memset(&r, 0, sizeof(r));
        int ids[3];
        for (int t = 0; t < 3; ++t) {
            // reference boundary test: (cellsOnVertex - 1 as size_t) > nCells + 1  (ST:33-40); ids in
            // (nCells-1, nCells+1] would index out of range in the reference -- treated as boundary here
            const long long e = (long long)cells_on_vertex[3 * ve + t] - 1;
            if (e < 0 || e >= (long long)n_cells) { r.boundary = 1; ids[t] = 0; }
            else ids[t] = (int)e;
            h_vcell[3 * i + t] = ids[t];
        }
        r.c0 = c_ext2int[ids[0]]; r.c1 = c_ext2int[ids[1]]; r.c2 = c_ext2int[ids[2]];
        h_vert[i] = r;
    }
    double *d_cell_ext = nullptr, *d_vert_ext = nullptr;
    CK(cudaMalloc(&ctx->c4, nC * sizeof(double4)));
    CK(cudaMalloc(&ctx->trig, nC * sizeof(double4)));
    CK(cudaMalloc(&ctx->vert, nV * sizeof(VertRec)));
    CK(cudaMalloc(&ctx->vcell_ext, nV * 3 * sizeof(int)));
    CK(cudaMalloc(&ctx->c_int2ext, nC * sizeof(int)));
    CK(cudaMalloc(&ctx->c_ext2int, nC * sizeof(int)));
    CK(cudaMalloc(&ctx->v_int2ext, nV * sizeof(int)));
    CK(cudaMalloc(&ctx->v_ext2int, nV * sizeof(int)));
    CK(cudaMalloc(&d_cell_ext, nC * 24));
    CK(cudaMalloc(&d_vert_ext, nV * 24));
    ctx->mesh_bytes += nC * (2 * sizeof(double4) + 8) + nV * (sizeof(VertRec) + 12 + 8);
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->c4, h_c4.data(), nC * sizeof(double4), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->vert, h_vert.data(), nV * sizeof(VertRec), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->vcell_ext, h_vcell.data(), nV * 3 * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->c_int2ext, c_int2ext.data(), nC * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->c_ext2int, c_ext2int.data(), nC * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->v_int2ext, v_int2ext.data(), nV * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->v_ext2int, v_ext2int.data(), nV * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cell_ext, cell_xyz, nC * 24, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_vert_ext, vertex_xyz, nV * 24, cudaMemcpyHostToDevice, st));
    k_cell_trig<<<blocks_for(n_cells, 256), 256, 0, st>>>(d_cell_ext, ctx->trig, n_cells);
    k_vert_bary<<<blocks_for(n_vertices, 256), 256, 0, st>>>(ctx->vert, ctx->v_int2ext, d_vert_ext, d_cell_ext, ctx->vcell_ext, n_vertices);
    ctx->launches += 2;
    CK(cudaGetLastError());
    switch (ctx->M) {
    case 6: rc = build_cube<6>(ctx); break;
    case 8: rc = build_cube<8>(ctx); break;
    default: rc = build_cube<20>(ctx); break;
    }
    if (rc) return rc;
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d_cell_ext));
    CK(cudaFree(d_vert_ext));

    // keep the mesh records resident in L2 where the device allows it (access-policy window on
    // both streams; hitRatio scaled when the records exceed the persisting carve-out)
    if (ctx->prop.persistingL2CacheMaxSize > 0 && ctx->prop.accessPolicyMaxWindowSize > 0 && !getenv("MOPS_NO_L2_WINDOW")) {
        const size_t rec_bytes = (ctx->M == 6 ? sizeof(CellRec<6>) : ctx->M == 8 ? sizeof(CellRec<8>) : sizeof(CellRec<20>)) * nC;
        const size_t carve = std::min<size_t>((size_t)ctx->prop.persistingL2CacheMaxSize, rec_bytes);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
            cudaStreamAttrValue attr;
            std::memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = ctx->rec;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(rec_bytes, (size_t)ctx->prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)attr.accessPolicyWindow.num_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            ctx->l2_attr = attr;
            ctx->l2_attr_valid = true;
        }
        cudaGetLastError();
    }
    ctx->has_mesh = true;
    return MOPS_OK;
}

int mops_set_snapshot(mops_ctx* ctx, int32_t slot, int32_t n_levels, const double* zonal, const double* meridional,
                      const double* layer_thickness, const double* bottom_depth, const double* vert_vel_top, int32_t n_attr,
                      const double* const* attrs, int32_t n_attr_total)
{
    return set_snapshot_impl(ctx, slot, n_levels, zonal, meridional, layer_thickness, bottom_depth, vert_vel_top, n_attr, attrs,
                             n_attr_total, false);
}

int mops_set_snapshot_async(mops_ctx* ctx, int32_t slot, int32_t n_levels, const double* zonal, const double* meridional,
                            const double* layer_thickness, const double* bottom_depth, const double* vert_vel_top, int32_t n_attr,
                            const double* const* attrs, int32_t n_attr_total)
{
    return set_snapshot_impl(ctx, slot, n_levels, zonal, meridional, layer_thickness, bottom_depth, vert_vel_top, n_attr, attrs,
                             n_attr_total, true);
}

int mops_side_wait_event(mops_ctx* ctx, void* cuda_event)
{
    if (!ctx || !cuda_event) return MOPS_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamWaitEvent(ctx->side, (cudaEvent_t)cuda_event, 0));
    return MOPS_OK;
}

int mops_snapshot_wait(mops_ctx* ctx, int32_t slot)
{
    if (!ctx) return MOPS_E_INVALID;
    if (slot < 0 || slot >= MOPS_MAX_SNAPSHOT_SLOTS || !ctx->snap[slot].valid) return fail(ctx, MOPS_E_STATE, "slot %d not set", slot);
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->snap[slot].ready));
    return MOPS_OK;
}

int mops_get_prepared(mops_ctx* ctx, int32_t slot, double* ztop_vertex, double* vel_vertex, double* vertvel_vertex,
                      double* attr0_vertex, double* attr1_vertex)
{
    if (!ctx) return MOPS_E_INVALID;
    if (slot < 0 || slot >= MOPS_MAX_SNAPSHOT_SLOTS || !ctx->snap[slot].valid) return fail(ctx, MOPS_E_STATE, "slot %d not set", slot);
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = wait_slot(ctx, slot))) return rc;
    Snapshot& s = ctx->snap[slot];
    const size_t nV = (size_t)ctx->nV, L = (size_t)s.L;
    double *d_z = nullptr, *d_v = nullptr, *d_w = nullptr, *d_a0 = nullptr, *d_a1 = nullptr;
    if (ztop_vertex) CK(cudaMalloc(&d_z, nV * L * 8));
    if (vel_vertex) CK(cudaMalloc(&d_v, nV * L * 24));
    if (vertvel_vertex) CK(cudaMalloc(&d_w, nV * (L + 1) * 8));
    if (attr0_vertex && s.attr[0]) CK(cudaMalloc(&d_a0, nV * L * 8));
    if (attr1_vertex && s.attr[1]) CK(cudaMalloc(&d_a1, nV * L * 8));
    k_export_prepared<<<blocks_for((long long)nV * L, 256), 256, 0, ctx->stream>>>(ctx->v_ext2int, s.ztop, s.velw, s.attr[0], s.attr[1],
                                                                                 d_z, d_v, d_w, d_a0, d_a1, ctx->nV, s.L, s.w_is_z ? 1 : 0);
    ctx->launches++;
    CK(cudaGetLastError());
    if (d_z) CK(cudaMemcpyAsync(ztop_vertex, d_z, nV * L * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_v) CK(cudaMemcpyAsync(vel_vertex, d_v, nV * L * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_w) CK(cudaMemcpyAsync(vertvel_vertex, d_w, nV * (L + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_a0) CK(cudaMemcpyAsync(attr0_vertex, d_a0, nV * L * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_a1) CK(cudaMemcpyAsync(attr1_vertex, d_a1, nV * L * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_z); cudaFree(d_v); cudaFree(d_w); cudaFree(d_a0); cudaFree(d_a1);
    return MOPS_OK;
}

int mops_locate(mops_ctx* ctx, int32_t mem, int64_t n, const double* xyz, int32_t* cell_out)
{
    if (!ctx) return MOPS_E_INVALID;
    if (!ctx->has_mesh) return fail(ctx, MOPS_E_STATE, "no mesh");
    if (n < 0 || (n > 0 && (!xyz || !cell_out))) return fail(ctx, MOPS_E_INVALID, "bad locate arguments");
    if (n == 0) return MOPS_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if (mem == MOPS_MEM_HOST) {
        if ((rc = ensure(ctx, ctx->p_xyz, (size_t)n * 24))) return rc;
        if ((rc = ensure(ctx, ctx->p_cell0, (size_t)n * 4))) return rc;
        CK(cudaMemcpyAsync(ctx->p_xyz.p, xyz, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream));
        dispatch_locate(ctx, n, (const double*)ctx->p_xyz.p, nullptr, (int*)ctx->p_cell0.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(cell_out, ctx->p_cell0.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        dispatch_locate(ctx, n, xyz, nullptr, cell_out);
        CK(cudaGetLastError());
    }
    return MOPS_OK;
}

int mops_order_key(mops_ctx* ctx, int32_t mem, int64_t n, const double* xyz, int32_t* key_out)
{
    if (!ctx) return MOPS_E_INVALID;
    if (!ctx->has_mesh) return fail(ctx, MOPS_E_STATE, "no mesh");
    if (n < 0 || (n > 0 && (!xyz || !key_out))) return fail(ctx, MOPS_E_INVALID, "bad order-key arguments");
    if (n == 0) return MOPS_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    // the engine numbers its cells along a Morton curve, so the internal id of a point's cell IS its rank on the curve
    if (mem == MOPS_MEM_HOST) {
        if ((rc = ensure(ctx, ctx->p_xyz, (size_t)n * 24))) return rc;
        if ((rc = ensure(ctx, ctx->p_cell0, (size_t)n * 4))) return rc;
        CK(cudaMemcpyAsync(ctx->p_xyz.p, xyz, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream));
        dispatch_locate(ctx, n, (const double*)ctx->p_xyz.p, (int*)ctx->p_cell0.p, nullptr);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(key_out, ctx->p_cell0.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        dispatch_locate(ctx, n, xyz, key_out, nullptr);
        CK(cudaGetLastError());
    }
    return MOPS_OK;
}

void* mops_get_stream(mops_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int mops_get_device(mops_ctx* ctx) { return ctx ? ctx->device : -1; }

int mops_streamline(mops_ctx* ctx, const mops_traj_cfg* cfg, int32_t slot, const mops_traj_io* io, mops_traj_stats* stats)
{
    return trajectory_impl(ctx, cfg, slot, slot, io, stats, false, nullptr);
}

int mops_pathline(mops_ctx* ctx, const mops_traj_cfg* cfg, int32_t front_slot, int32_t back_slot, const mops_traj_io* io,
                  mops_traj_stats* stats)
{
    return trajectory_impl(ctx, cfg, front_slot, back_slot, io, stats, true, nullptr);
}

int mops_streamline_submit(mops_ctx* ctx, const mops_traj_cfg* cfg, int32_t slot, const mops_traj_io* io, int64_t* ticket)
{
    if (!ticket) return MOPS_E_INVALID;
    long long tk = 0;
    const int rc = trajectory_impl(ctx, cfg, slot, slot, io, nullptr, false, &tk);
    *ticket = tk;
    return rc;
}

int mops_pathline_submit(mops_ctx* ctx, const mops_traj_cfg* cfg, int32_t front_slot, int32_t back_slot, const mops_traj_io* io,
                         int64_t* ticket)
{
    if (!ticket) return MOPS_E_INVALID;
    long long tk = 0;
    const int rc = trajectory_impl(ctx, cfg, front_slot, back_slot, io, nullptr, true, &tk);
    *ticket = tk;
    return rc;
}

int mops_traj_wait(mops_ctx* ctx, int64_t ticket, int32_t what, mops_traj_stats* stats)
{
    if (!ctx) return MOPS_E_INVALID;
    if (ticket == 0) { // n == 0 submit: nothing was enqueued
        if (stats) std::memset(stats, 0, sizeof(*stats));
        return MOPS_OK;
    }
    return host_wait(ctx, ticket, what, stats);
}

int mops_remap_fixed_depth(mops_ctx* ctx, const mops_remap_cfg* cfg, int32_t slot, double* img0, double* img1, int32_t* pixel_cell,
                           mops_remap_stats* stats)
{
    if (!ctx) return MOPS_E_INVALID;
    if (!cfg || !img0) return fail(ctx, MOPS_E_INVALID, "null cfg/img0");
    if (!ctx->has_mesh) return fail(ctx, MOPS_E_STATE, "no mesh");
    if (slot < 0 || slot >= MOPS_MAX_SNAPSHOT_SLOTS || !ctx->snap[slot].valid) return fail(ctx, MOPS_E_STATE, "snapshot slot %d not set", slot);
    if (cfg->width <= 0 || cfg->height <= 0) return fail(ctx, MOPS_E_INVALID, "bad image size");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = wait_slot(ctx, slot))) return rc;
    Snapshot& S = ctx->snap[slot];
    const size_t npx = (size_t)cfg->width * cfg->height;
    const bool host = (cfg->mem == MOPS_MEM_HOST);
    const bool attr_image = (S.n_attr_total > 1) && img1; // VK:259-267
    const long long launches0 = ctx->launches;
    cudaStream_t st = ctx->stream;
    double *d0, *d1 = nullptr;
    int* dc = nullptr;
    CK(cudaEventRecord(ctx->ev0, st));
    if (host) {
        if ((rc = ensure(ctx, ctx->r_img0, npx * 32))) return rc;
        d0 = (double*)ctx->r_img0.p;
        if (attr_image) { if ((rc = ensure(ctx, ctx->r_img1, npx * 32))) return rc; d1 = (double*)ctx->r_img1.p; }
        if (pixel_cell) { if ((rc = ensure(ctx, ctx->r_cells, npx * 4))) return rc; dc = (int*)ctx->r_cells.p; }
    } else {
        d0 = img0; d1 = attr_image ? img1 : nullptr; dc = pixel_cell;
    }
    CK(cudaMemsetAsync(ctx->counters, 0, 8 * sizeof(unsigned long long), st));
    RemapParams P;
    std::memset(&P, 0, sizeof(P));
    P.rec = ctx->rec; P.c4 = ctx->c4; P.cube = ctx->cube; P.c_int2ext = ctx->c_int2ext; P.F = ctx->F; P.nC = ctx->nC; P.L = S.L;
    P.kd = kd_view(ctx);
    P.s = view_of(S);
    P.attr_count = S.n_attr; P.attr_image = attr_image ? 1 : 0;
    P.width = cfg->width; P.height = cfg->height;
    P.minLat = cfg->lat_min; P.maxLat = cfg->lat_max; P.minLon = cfg->lon_min; P.maxLon = cfg->lon_max;
    P.DEPTH = -cfg->fixed_depth;
    P.img0 = d0; P.img1 = d1; P.pixel_cell = dc; P.nan_count = ctx->counters + 2;
    CK(cudaEventRecord(ctx->ev1, st));
    switch (ctx->M) {
    case 6: launch_remap<6>(ctx, P); break;
    case 8: launch_remap<8>(ctx, P); break;
    default: launch_remap<20>(ctx, P); break;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev2, st));
    if ((rc = mark_use(ctx, slot))) return rc;
    unsigned long long h_counters[4] = {0, 0, 0, 0};
    if (host) {
        CK(cudaMemcpyAsync(img0, d0, npx * 32, cudaMemcpyDeviceToHost, st));
        if (d1) CK(cudaMemcpyAsync(img1, d1, npx * 32, cudaMemcpyDeviceToHost, st));
        if (dc) CK(cudaMemcpyAsync(pixel_cell, dc, npx * 4, cudaMemcpyDeviceToHost, st));
    }
    if (host || stats) {
        CK(cudaMemcpyAsync(h_counters, ctx->counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev3, st));
        CK(cudaStreamSynchronize(st));
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev2) == cudaSuccess) stats->kernel_ms = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev3) == cudaSuccess) stats->total_ms = ms;
        stats->nan_pixels = (int64_t)h_counters[2];
        stats->launches = (int32_t)(ctx->launches - launches0);
        stats->n_images = attr_image ? 2 : 1;
    }
    return MOPS_OK;
}

// Line assembly + NaN trimming exactly as the reference's host code does it
// (InitTrajectoryLines / FinalizeTrajectoryLines[WithAttrs] / RemoveNaNTrajectoriesAndReindex,
// src/Common/TrajectoryCommon.h:43-190; pinned by the reference's test/test_trajector.cpp).
int mops_finalize_lines(int64_t n, int32_t each, const double* seeds, const double* raw_pos, const double* raw_vel,
                        int32_t pathline_mode, double* points, double* velocity, double* temperature, double* salinity,
                        double* last)
{
    if (n < 0 || each <= 0 || (n > 0 && (!seeds || !raw_pos || !raw_vel || !points || !velocity))) return MOPS_E_INVALID;
    const int64_t per = (int64_t)each + 1;
    auto finite = [](const double* p) { return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); };
    // lines are independent: large calls are split over the host cores (1 M lines x 169 points is ~2 s on one core, which
    // would dwarf the kernels that produced them)
    auto run = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
        double* P = points + i * per * 3;
        double* V = velocity + i * per * 3;
        double* T = temperature ? temperature + i * per : nullptr;
        double* S = salinity ? salinity + i * per : nullptr;
        // points = [seed, rec_0 .. rec_each-1]; velocity = [vel_0 .. vel_each-1, 0] (one shorter, zero padded)
        std::memcpy(P, seeds + 3 * i, 24);
        std::memcpy(P + 3, raw_pos + i * each * 3, (size_t)each * 24);
        std::memcpy(V, raw_vel + i * each * 3, (size_t)each * 24);
        V[3 * each] = V[3 * each + 1] = V[3 * each + 2] = 0.0;
        for (int64_t k = 0; k < per; ++k) {
            // FinalizeTrajectoryLinesWithAttrs pushes velocity.x / velocity.y, not the attributes (:179-180)
            const bool src = pathline_mode && k < each;
            if (T) T[k] = src ? raw_vel[(i * each + k) * 3] : 0.0;
            if (S) S[k] = src ? raw_vel[(i * each + k) * 3 + 1] : 0.0;
        }
        int64_t cut = 0;
        for (; cut < per; ++cut)
            if (!finite(P + 3 * cut)) break;
        if (cut == 0) {
            const double ft = T ? T[0] : 0.0, fs = S ? S[0] : 0.0;
            for (int64_t j = 0; j < per; ++j) {
                std::memcpy(P + 3 * j, P, 24);
                V[3 * j] = V[3 * j + 1] = V[3 * j + 2] = 0.0;
                if (T) T[j] = ft;
                if (S) S[j] = fs;
            }
        } else if (cut < per) {
            const double lt = T ? T[cut - 1] : 0.0, ls = S ? S[cut - 1] : 0.0;
            V[3 * (cut - 1)] = V[3 * (cut - 1) + 1] = V[3 * (cut - 1) + 2] = 0.0;
            for (int64_t j = cut; j < per; ++j) {
                std::memcpy(P + 3 * j, P + 3 * (cut - 1), 24);
                V[3 * j] = V[3 * j + 1] = V[3 * j + 2] = 0.0;
                if (T) T[j] = lt;
                if (S) S[j] = ls;
            }
        }
        if (last) std::memcpy(last + 3 * i, P + 3 * (per - 1), 24);
    }
    };
    const int64_t min_chunk = std::max<int64_t>(1, (1 << 20) / per); // ~1 M points per task at least
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = getenv("MOPS_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(e));
    nt = (unsigned)std::min<int64_t>(std::max(1u, std::min(nt, 64u)), (n + min_chunk - 1) / min_chunk);
    if (nt <= 1) {
        run(0, n);
    } else {
        std::vector<std::thread> pool;
        const int64_t chunk = (n + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t) {
            const int64_t lo = (int64_t)t * chunk, hi = std::min(n, lo + chunk);
            if (lo < hi) pool.emplace_back(run, lo, hi);
        }
        for (auto& th : pool) th.join();
    }
    return MOPS_OK;
}

int mops_remap_fixed_layer(mops_ctx* ctx, const mops_view_cfg* cfg, int32_t slot, double* img, int32_t* pixel_cell,
                           mops_remap_stats* stats)
{
    return view_impl(ctx, cfg, slot, img, pixel_cell, stats, 0);
}

int mops_regrid_fixed_latitude(mops_ctx* ctx, const mops_view_cfg* cfg, int32_t slot, double* img, int32_t* pixel_cell,
                               mops_remap_stats* stats)
{
    return view_impl(ctx, cfg, slot, img, pixel_cell, stats, 1);
}

int mops_get_info(mops_ctx* ctx, mops_info* out)
{
    if (!ctx || !out) return MOPS_E_INVALID;
    std::memset(out, 0, sizeof(*out));
    out->device = ctx->device;
    out->sm_count = ctx->prop.multiProcessorCount;
    out->cc_major = ctx->prop.major;
    out->cc_minor = ctx->prop.minor;
    out->l2_bytes = ctx->prop.l2CacheSize;
    out->hbm_bytes = (int64_t)ctx->prop.totalGlobalMem;
    out->mesh_bytes = (int64_t)ctx->mesh_bytes;
    for (int i = 0; i < MOPS_MAX_SNAPSHOT_SLOTS; ++i) {
        out->snapshot_bytes[i] = (int64_t)ctx->snap[i].bytes;
        if (ctx->snap[i].valid && ctx->snap[i].nonmono < 0 && cudaEventQuery(ctx->snap[i].ready) == cudaSuccess) {
            // async upload: fetch the count lazily once its preprocessing has completed
            cudaMemcpy(&ctx->snap[i].nonmono, ctx->d_nonmono + i, sizeof(int), cudaMemcpyDeviceToHost);
        }
        out->nonmonotone_cells[i] = ctx->snap[i].nonmono;
        if (ctx->snap[i].valid) out->n_levels = ctx->snap[i].L;
    }
    out->record_width = ctx->M;
    out->total_launches = ctx->launches.load();
    return MOPS_OK;
}

} // extern "C"
