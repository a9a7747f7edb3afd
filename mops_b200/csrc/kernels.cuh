// kernels.cuh -- every __global__ kernel of the engine (sm_100a).  Host-side launch logic is
// in engine.cu.  VK = src/CPU/TBB/Kernel/MPASOVisualizerKernels.cpp, TK = .../TBBKernel.h,
// ST = src/CPU/TBB/MPASOSolutionTBB.cpp of the reference.
#pragma once
#include "engine.cuh"
#include "fastpath.cuh"

namespace mops {

// =========================================================================================
// mesh set-up kernels (run once per mesh)
// =========================================================================================

// point-independent parts of IsInMesh (TK:40-47) and of Wachspress (Interpolation.hpp:154)
template <int M>
__global__ void k_build_records(CellRec<M>* __restrict__ rec, int nC)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    CellRec<M>* r = rec + c;
    const int nv = r->nv;
    for (int k = 0; k < M; ++k) {
        if (k < nv) {
            const int kn = (k + 1) % nv;
            const int kp = (k - 1 + nv) % nv;
            const double ax = r->vx[k], ay = r->vy[k], az = r->vz[k];
            const double bx = r->vx[kn], by = r->vy[kn], bz = r->vz[kn];
            r->nx[k] = ay * bz - az * by; // cy::Vec3::Cross
            r->ny[k] = az * bx - ax * bz;
            r->nz[k] = ax * by - ay * bx;
            r->B[k] = tri_area(r->vx[kp], r->vy[kp], r->vz[kp], ax, ay, az, bx, by, bz);
        } else {
            r->nx[k] = 0.0; r->ny[k] = 0.0; r->nz[k] = 0.0; r->B[k] = 0.0;
        }
    }

}

// coefficients of GeoConverter::convertENUVelocityToXYZ (GeoConverter.hpp:225-250) at a cell
// centre, stored in the CALLER's cell order: (slon, clon, slat, clat); slon = NaN flags the
// x == y == 0 singularity branch.
__global__ void k_cell_trig(const double* __restrict__ cell_xyz_ext, double4* __restrict__ trig, int nC)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const double x = cell_xyz_ext[3 * (size_t)c], y = cell_xyz_ext[3 * (size_t)c + 1], z = cell_xyz_ext[3 * (size_t)c + 2];
    double4 t;
    if (x == 0.0 && y == 0.0) {
        t.x = nan(""); t.y = 0.0; t.z = 0.0; t.w = 0.0;
    } else {
        const double Rxy = sqrt(x * x + y * y);
        const double Rxyz = sqrt(x * x + y * y + z * z);
        t.x = y / Rxy;    // slon
        t.y = x / Rxy;    // clon
        t.z = z / Rxyz;   // slat
        t.w = Rxy / Rxyz; // clat
    }
    trig[c] = t;
}

// Interpolator::calcTriangleBarycentric of each Voronoi vertex in the triangle of its three
// cell centres (Interpolation.hpp:79-93); mesh-constant, so computed once instead of once per
// (vertex, level) as ST:42-52 does.
__global__ void k_vert_bary(VertRec* __restrict__ vert, const int* __restrict__ vext, const double* __restrict__ vertex_xyz_ext,
                            const double* __restrict__ cell_xyz_ext, const int* __restrict__ vcell_ext, int nV)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    VertRec r = vert[v];
    r.u = 0.0; r.v = 0.0; r.w = 0.0;
    if (!r.boundary) {
        const size_t ve = (size_t)vext[v];
        const double px = vertex_xyz_ext[3 * ve], py = vertex_xyz_ext[3 * ve + 1], pz = vertex_xyz_ext[3 * ve + 2];
        const size_t e0 = (size_t)vcell_ext[3 * (size_t)v], e1 = (size_t)vcell_ext[3 * (size_t)v + 1], e2 = (size_t)vcell_ext[3 * (size_t)v + 2];
        const double t0x = cell_xyz_ext[3 * e0], t0y = cell_xyz_ext[3 * e0 + 1], t0z = cell_xyz_ext[3 * e0 + 2];
        const double v0x = cell_xyz_ext[3 * e1] - t0x, v0y = cell_xyz_ext[3 * e1 + 1] - t0y, v0z = cell_xyz_ext[3 * e1 + 2] - t0z;
        const double v1x = cell_xyz_ext[3 * e2] - t0x, v1y = cell_xyz_ext[3 * e2 + 1] - t0y, v1z = cell_xyz_ext[3 * e2 + 2] - t0z;
        const double v2x = px - t0x, v2y = py - t0y, v2z = pz - t0z;
        const double d00 = v0x * v0x + v0y * v0y + v0z * v0z;
        const double d01 = v0x * v1x + v0y * v1y + v0z * v1z;
        const double d11 = v1x * v1x + v1y * v1y + v1z * v1z;
        const double d20 = v2x * v0x + v2y * v0y + v2z * v0z;
        const double d21 = v2x * v1x + v2y * v1y + v2z * v1z;
        const double denom = d00 * d11 - d01 * d01;
        r.v = (d11 * d20 - d01 * d21) / denom;
        r.w = (d00 * d21 - d01 * d20) / denom;
        r.u = 1.0 - r.v - r.w;
    }
    vert[v] = r;
}

// cube-map start table of the locate walk, built coarse-to-fine: bucket (face,i,j) at
// resolution F starts its walk from the parent bucket's answer at resolution F/2.
template <int M>
__global__ void k_cube_level(const CellRec<M>* __restrict__ rec, const double4* __restrict__ c4, const int* __restrict__ parent,
                             int* __restrict__ table, int F, double radius)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 6 * F * F) return;
    const int face = id / (F * F);
    const int i = (id / F) % F;
    const int j = id % F;
    double x, y, z;
    cube_center(face, i, j, F, x, y, z);
    const double s = radius / sqrt(x * x + y * y + z * z);
    const int start = parent ? parent[(face * (F / 2) + i / 2) * (F / 2) + j / 2] : 0;
    table[id] = walk_nearest<M>(rec, c4, start, x * s, y * s, z * s);
}

// =========================================================================================
// point location: 8 lanes cooperate on one query (one lane per cellsOnCell neighbour)
// replaces MPASOField::calcInWhichCells (src/Core/MPASOField.cpp:23-34)
// =========================================================================================
template <int M>
__global__ void __launch_bounds__(256) k_locate(const CellRec<M>* __restrict__ rec, const double4* __restrict__ c4,
                                                const int* __restrict__ cube, int F, int nC, const KdView kd,
                                                long long n, const double* __restrict__ xyz,
                                                int* __restrict__ cell_int, int* __restrict__ cell_ext,
                                                const int* __restrict__ c_int2ext)
{
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    if (gid >= n) return; // whole 8-lane group exits together
    const double qx = xyz[3 * gid], qy = xyz[3 * gid + 1], qz = xyz[3 * gid + 2];
    int result = -1;
    if (finite3(qx, qy, qz)) {
        int cur = cube[cube_bucket(qx, qy, qz, F)];
        double4 c = c4[cur];
        double dcur = dist2(qx, qy, qz, c.x, c.y, c.z);
        for (int it = 0; it < (1 << 22); ++it) {
            const CellRec<M>* r = rec + cur;
            const int nv = r->nv;
            double dbest = dcur;
            int best = cur;
            for (int base = 0; base < nv; base += 8) { // one pass for nv <= 8
                const int k = base + sub;
                if (k < nv) {
                    const int nb = r->nbr[k];
                    if (nb >= 0) {
                        const double4 cc = c4[nb];
                        const double d = dist2(qx, qy, qz, cc.x, cc.y, cc.z);
                        if (d < dbest) { dbest = d; best = nb; }
                    }
                }
            }
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) { // argmin over the 8 lanes (ties -> lower cell id)
                const double od = __shfl_xor_sync(gmask, dbest, off, 8);
                const int ob = __shfl_xor_sync(gmask, best, off, 8);
                if (od < dbest || (od == dbest && ob < best)) { dbest = od; best = ob; }
            }
            if (!(dbest < dcur)) break;
            cur = best;
            dcur = dbest;
        }
        result = cur;
        // culled mesh: accept only a cell that contains the query, else exact kd search (lane 0 decides)
        if (kd.n > 0 && sub == 0 && !in_mesh<M>(rec + cur, rec[cur].nv, qx, qy, qz)) result = kd_nearest(kd, qx, qy, qz, cur, dcur);
    }
    if (sub == 0) {
        if (cell_int) cell_int[gid] = result;
        if (cell_ext) cell_ext[gid] = result >= 0 ? c_int2ext[result] : -1;
    }
}

// caller cell ids -> internal ids (and back)
__global__ void k_map_ids(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ map, int nmap, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = in[i];
    out[i] = (c >= 0 && c < nmap) ? map[c] : -1;
}

__global__ void k_iota(int* __restrict__ a, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}

// =========================================================================================
// snapshot preprocessing (a17) -- replaces MOPSApp::addSol's chain, src/Core/MOPSApp.cpp:100-130
// =========================================================================================

// MPASOSolution::calcCellCenterZtop, bottomDepth branch (src/Core/MPASOSolution.cpp:565-577): zTop[c][k] = -bottomDepth[c] +
// sum_{j >= k} thickness[c][j], summed bottom-up in exactly that order (the association decides the bits), so one thread
// walks one cell's column.  The column rows are staged through shared memory by the whole block -- global loads and stores
// are coalesced along k, the serial walk runs on the staged copy (row stride L + 1 doubles: conflict-free) -- which took the
// kernel from 11.6 ms to the time of its 2 x 1.7 GB of traffic on the level-9 x 80-layer snapshot.
constexpr int CZ_CELLS = 32; // cells per block: 32 x 101 doubles = 26 KB of shared memory at the largest L (100)
__global__ void __launch_bounds__(128) k_cell_ztop(const double* __restrict__ thick, const double* __restrict__ bottom, double* __restrict__ ztop_c,
                                                   int nC, int L)
{
    extern __shared__ double cz_tile[]; // [CZ_CELLS][L + 1]
    const int c0 = blockIdx.x * CZ_CELLS;
    const int nc = min(CZ_CELLS, nC - c0);
    const int ld = L + 1;
    const size_t base = (size_t)c0 * L;
    for (int i = threadIdx.x; i < nc * L; i += blockDim.x) cz_tile[(i / L) * ld + (i % L)] = thick[base + i];
    __syncthreads();
    if (threadIdx.x < nc) {
        double* row = cz_tile + threadIdx.x * ld;
        double z = -bottom[c0 + threadIdx.x];
        for (int k = L - 1; k >= 0; --k) {
            z += row[k];
            row[k] = z * 1.0;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nc * L; i += blockDim.x) ztop_c[base + i] = cz_tile[(i / L) * ld + (i % L)];
}

// CalcCellVertexZtop (ST:9-55) + CalcCellCenterVelocityByZM (ST:108-129) + CalcCellVertexVelocity
// (ST:270-318) + CalcCellVertexVertVelocity (ST:320-366) fused: one thread per (vertex, level),
// level fastest so reads of the three cell rows and the writes are coalesced along k.
__global__ void k_vertex_fields(const VertRec* __restrict__ vert, const int* __restrict__ vcell_ext, const double4* __restrict__ trig,
                                const double* __restrict__ ztop_c, const double* __restrict__ zonal, const double* __restrict__ merid,
                                const double* __restrict__ wtop, double* __restrict__ ztop_v, double4* __restrict__ velw_v,
                                int nV, int L, int* __restrict__ any_w, int pack_z)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nV * L) return;
    const int v = (int)(idx / L);
    const int k = (int)(idx % L);
    const VertRec r = vert[v];
    double zt = 0.0;
    double4 o = make_double4(0.0, 0.0, 0.0, 0.0);
    if (!r.boundary) {
        double zc[3], wc[3], vx[3], vy[3], vz[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const size_t e = (size_t)vcell_ext[3 * (size_t)v + t];
            zc[t] = ztop_c[e * L + k];
            wc[t] = wtop ? wtop[e * (L + 1) + k] : 0.0;
            const double Uzon = zonal[e * L + k], Umer = merid[e * L + k], Uup = 0.0;
            const double4 tr = trig[e];
            if (isnan(tr.x)) {
                vx[t] = 0.0; vy[t] = 0.0; vz[t] = Uup;
            } else {
                const double slon = tr.x, clon = tr.y, slat = tr.z, clat = tr.w;
                vx[t] = -slon * Uzon - slat * clon * Umer + clon * clat * Uup;
                vy[t] = clon * Uzon - slat * slon * Umer + slon * clat * Uup;
                vz[t] = clat * Umer + slat * Uup;
            }
        }
        zt = r.u * zc[0] + r.v * zc[1] + r.w * zc[2];
        o.x = vx[0] * r.u + vx[1] * r.v + vx[2] * r.w;
        o.y = vy[0] * r.u + vy[1] * r.v + vy[2] * r.w;
        o.z = vz[0] * r.u + vz[1] * r.v + vz[2] * r.w;
        o.w = r.u * wc[0] + r.v * wc[1] + r.w * wc[2];
    }
    ztop_v[idx] = zt;
    // does the snapshot carry any vertical velocity?  (bit test: -0.0 and NaN count as "yes")
    if (__double_as_longlong(o.w) != 0ll) *any_w = 1;
    if (pack_z) o.w = zt; // no vertVelocityTop was given: the w slot carries zTop (SnapView::w_is_z)
    velw_v[idx] = o;
}

// CalcCellCenterToVertex (ST:57-106): scalar attribute, clamped >= 0
__global__ void k_vertex_scalar(const VertRec* __restrict__ vert, const int* __restrict__ vcell_ext, const double* __restrict__ cell_val,
                                double* __restrict__ vert_val, int nV, int L)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nV * L) return;
    const int v = (int)(idx / L);
    const int k = (int)(idx % L);
    const VertRec r = vert[v];
    double out = 0.0;
    if (!r.boundary) {
        const double a = cell_val[(size_t)vcell_ext[3 * (size_t)v] * L + k];
        const double b = cell_val[(size_t)vcell_ext[3 * (size_t)v + 1] * L + k];
        const double c = cell_val[(size_t)vcell_ext[3 * (size_t)v + 2] * L + k];
        out = r.u * a + r.v * b + r.w * c;
        if (out < 0.0) out = 0.0;
    }
    vert_val[idx] = out;
}

// per-vertex: is the zTop column finite and non-increasing?  (one warp per vertex)
__global__ void k_vertex_mono(const double* __restrict__ ztop_v, unsigned char* __restrict__ vmono, int nV, int L)
{
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= nV) return;
    const double* col = ztop_v + (size_t)gw * L;
    bool ok = true;
    for (int k = lane; k < L; k += 32) {
        const double a = col[k];
        if (!isfinite(a)) ok = false;
        if (k > 0 && a > col[k - 1]) ok = false;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) vmono[gw] = ok ? 1 : 0;
}

template <int M>
__global__ void k_cell_mono(const CellRec<M>* __restrict__ rec, const unsigned char* __restrict__ vmono, unsigned char* __restrict__ cmono,
                            int nC, int* __restrict__ n_nonmono)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const CellRec<M>* r = rec + c;
    bool ok = true;
    for (int k = 0; k < r->nv; ++k) ok = ok && (vmono[r->vid[k]] != 0);
    cmono[c] = ok ? 1 : 0;
    if (!ok) atomicAdd(n_nonmono, 1);
}

// vertex-major arrays back in the caller's vertex order (parity tests)
__global__ void k_export_prepared(const int* __restrict__ v_ext2int, const double* __restrict__ ztop_v, const double4* __restrict__ velw_v,
                                  const double* __restrict__ a0, const double* __restrict__ a1,
                                  double* __restrict__ o_ztop, double* __restrict__ o_vel, double* __restrict__ o_w,
                                  double* __restrict__ o_a0, double* __restrict__ o_a1, int nV, int L, int w_is_z)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nV * L) return;
    const int ve = (int)(idx / L);
    const int k = (int)(idx % L);
    const size_t src = (size_t)v_ext2int[ve] * L + k;
    if (o_ztop) o_ztop[idx] = ztop_v[src];
    const double4 q = velw_v[src];
    if (o_vel) { o_vel[3 * idx] = q.x; o_vel[3 * idx + 1] = q.y; o_vel[3 * idx + 2] = q.z; }
    if (o_w) {
        o_w[(size_t)ve * (L + 1) + k] = w_is_z ? 0.0 : q.w;
        if (k == L - 1) o_w[(size_t)ve * (L + 1) + L] = 0.0;
    }
    if (o_a0 && a0) o_a0[idx] = a0[src];
    if (o_a1 && a1) o_a1[idx] = a1[src];
}

// =========================================================================================
// streamline / pathline: one thread integrates one particle over all steps of the call
// (StreamLine VK:874-1003, PathLine VK:1329-1483)
// =========================================================================================
struct AdvectParams {
    const void* rec;
    const double4* c4;
    const int* c_int2ext;
    int nC, L;
    SnapView sv[2];   // [0] front, [1] back (pathline)
    int attr_count;   // pathline attributes in use (0..2)
    int no_w;         // host-side dispatch only (fastpath.cuh NOW): 0 = vertical velocity present, 1 = absent, 2 = absent and the w slots hold zTop
    int use_euler;
    int delta_t;      // signed seconds
    int times;        // steps
    int each;         // recorded slots per particle
    int record_t;     // seconds (streamline: run_time % record_t == 0)
    int record_interval; // pathline: (step+1) % (record_t/delta_t) == 0
    double duration;  // simulationDuration (pathline dalpha)
    long long n;
    const int* order; // processing order (particle indices sorted by start cell) or null
    double* pos;      // [n][3] in/out
    float* depth;     // [n]    in/out
    const int* cell0; // [n] internal start cells
    double* out_pos;  // [n][each][3] (zero-initialised by the host)
    double* out_vel;
    double* out_attr; // or null
    int* cell_log;    // [n][times] caller cell ids, or null
    int* status;      // or null
    int* steps;       // or null
    int* fcell;       // or null (caller cell ids)
    double* min_edge; // or null: per-particle smallest edge distance [rad] (diagnostic)
    int diag_edge;    // 1 = track the smallest edge distance of every evaluated point
    int walk;         // 1 = MOPS_SEM_WALK: evaluate each stage point in the cell that contains it
    unsigned long long* counters; // [0] particle-steps started, [1] alive at end, [3] near-edge particles, [4] stopped above the surface
    // segmented launches (SEG instantiations only): one launch integrates steps [step_begin, step_end) of the call;
    // particles still alive at step_end < times park their loop state in `state`, and the host compacts `order`
    // to the live ones before the next segment, so lanes of stopped particles do not ride along to the end.
    int step_begin, step_end;
    struct AdvState* state; // [n], indexed by particle
    const int* n_live;      // device-side particle count of this launch (the previous compaction's output) or null: P.n
};

struct alignas(32) AdvState {
    int cell, upd, started, hint_f, hint_b;
    int first_vel;
    int alive; // 1 = parked, to be resumed by the next segment
    int pad;
};

// cub::DeviceSelect::Flagged flag of position i of the processing order: inside the live prefix (whose length is on the
// device, so the host never reads it back and a segmented call stays asynchronous) and still alive
struct AdvAliveFlag {
    const int* order;
    const AdvState* state;
    const int* n_live; // null: every position is live (first compaction)
    __device__ __forceinline__ bool operator()(const int i) const
    {
        if (n_live && i >= *n_live) return false;
        return state[order[i]].alive != 0;
    }
};

__device__ __forceinline__ void st3(double* p, long long i, double x, double y, double z)
{
    p[3 * i] = x; p[3 * i + 1] = y; p[3 * i + 2] = z;
}


struct EvalPacked {
    EvalOut o;
    int st, hint_f, hint_b;
};
template <int M, bool PATH>
__device__ __noinline__ EvalPacked eval_generic_cold(const CellRec<M>* __restrict__ rec, const SnapView* __restrict__ sv, bool mf, bool mb,
                                                     int L, int attr_count, double px, double py, double pz, double depth, double alpha,
                                                     int hint_f, int hint_b)
{
    EvalPacked r;
    r.hint_f = hint_f; r.hint_b = hint_b;
    r.o.hx = 0.0; r.o.hy = 0.0; r.o.hz = 0.0; r.o.vv = 0.0; r.o.a0 = 0.0; r.o.a1 = 0.0;
    const d3 p = mk3(px, py, pz);
    r.st = PATH ? eval_path<M>(rec, sv, mf, mb, L, attr_count, p, depth, alpha, r.hint_f, r.hint_b, r.o)
                : eval_stream<M>(rec, sv[0], mf, L, p, depth, r.hint_f, r.o);
    return r;
}

// Block size / resident blocks of the advection kernel: build-time knobs so occupancy variants can be A/B'd
// (scripts/gpu_ab.sh); the defaults are the measured best on B200.
#ifndef MOPS_ADV_BLOCK
#define MOPS_ADV_BLOCK 128
#endif
#ifndef MOPS_ADV_MINB
#define MOPS_ADV_MINB 3
#endif

// state of one particle that a step reads and writes (by value in and out of the generic step, so that the caller's
// copies stay in registers when the generic step is an out-of-line call)
struct StepIO {
    d3 pos;           // in: start of step, out: end of step
    d3 hvel;          // out: the step's velocity (RK4: weighted mean of the four stages)
    double at0, at1;  // out: pathline attributes of the step
    double edge_min;  // in/out: smallest edge distance seen (EXTRA diagnostic)
    float depth_f;    // in/out
    int hint_f, hint_b;
    int status;       // out: ST_ALIVE or why the particle stopped (then nothing else is valid)
};

// One step of StreamLine / PathLine after the cell relocation, exactly as the reference does it (VK:923-986 / VK:1391-1465):
// Euler = one stage, RK4 = four, all stages against the start-of-step cell (R1); position, depth and radius update.
// This is the complete restatement: every instantiation without the straight-line fast path runs it inline, the
// production hexagon kernels call it out of line for the steps the fast path does not cover.
template <int M, bool PATH, bool EXTRA, bool ATTR>
__device__ __forceinline__ StepIO step_generic(const AdvectParams& P, int cell, int step, StepIO io)
{
    const CellRec<M>* __restrict__ recs = reinterpret_cast<const CellRec<M>*>(P.rec);
    const CellRec<M>* __restrict__ rec = recs + cell;
    const d3 pos = io.pos;
    const double dt = (double)P.delta_t;
    const double dalpha = PATH ? dt / P.duration : 0.0; // VK:1401
    const bool mono_f = P.sv[0].mono[cell] != 0;
    const bool mono_b = PATH ? (P.sv[1].mono[cell] != 0) : false;
    const double cur_depth = -1.0 * (double)io.depth_f;
    const double alpha = PATH ? (double)step / (double)P.times : 0.0; // VK:1345
    const double r = len3(pos);
    int hint_f = io.hint_f, hint_b = io.hint_b;
    d3 hvel = mk3(0.0, 0.0, 0.0);
    double vvel = 0.0, at0 = 0.0, at1 = 0.0;
    d3 new_pos;
    EvalOut o;

    // One rolled loop => one inlined copy of the evaluation.
    const int n_stage = P.use_euler ? 1 : 4;
    d3 hprev = mk3(0.0, 0.0, 0.0);
    int st = ST_ALIVE;
#pragma unroll 1
    for (int s = 0; s < n_stage; ++s) {
        d3 p = pos;
        double a_s = alpha;
        if (s > 0) {
            p = advect_on_sphere(pos, hprev, (s == 3) ? dt : dt * 0.5, r);
            if (PATH) a_s = clamp01(alpha + ((s == 3) ? dalpha : 0.5 * dalpha)); // VK:1410-1424
        }
        const CellRec<M>* __restrict__ rec_s = rec;
        bool mf = mono_f, mb = mono_b;
        if (EXTRA && P.walk && s > 0) { // MOPS_SEM_WALK: the cell that contains this stage point
            const int cs = walk_nearest<M>(recs, P.c4, cell, p.x, p.y, p.z);
            rec_s = recs + cs;
            mf = P.sv[0].mono[cs] != 0;
            mb = PATH ? (P.sv[1].mono[cs] != 0) : false;
        }
        if (EXTRA && P.diag_edge) {
            const double a = min_edge_angle<M>(rec_s, rec_s->nv, p.x, p.y, p.z);
            if (a < io.edge_min) io.edge_min = a;
        }
        if (M == 6) {
            // hexagon form of the evaluation (nv == M: no per-slot selects); whatever it does not cover (ST_GENERIC: the
            // 12 pentagons, coast cells, non-monotone columns, weight arithmetic outside the exact-sequence windows) takes
            // the general form out of line
            st = ST_GENERIC;
            if (rec_s->nv == M)
                st = PATH ? eval_path<M, true>(rec_s, P.sv, mf, mb, P.L, ATTR ? P.attr_count : 0, p, cur_depth, a_s, hint_f, hint_b, o)
                          : eval_stream<M, true>(rec_s, P.sv[0], mf, P.L, p, cur_depth, hint_f, o);
            if (st == ST_GENERIC) {
                const EvalPacked g = eval_generic_cold<M, PATH>(rec_s, P.sv, mf, mb, P.L, ATTR ? P.attr_count : 0, p.x, p.y, p.z,
                                                                cur_depth, a_s, hint_f, hint_b);
                st = g.st; hint_f = g.hint_f; hint_b = g.hint_b; o = g.o;
            }
        } else {
            st = PATH ? eval_path<M>(rec_s, P.sv, mf, mb, P.L, ATTR ? P.attr_count : 0, p, cur_depth, a_s, hint_f, hint_b, o)
                      : eval_stream<M>(rec_s, P.sv[0], mf, P.L, p, cur_depth, hint_f, o);
        }
        if (st != ST_ALIVE) break;
        if (s == 0) {
            hvel = mk3(o.hx, o.hy, o.hz);
            vvel = o.vv;
            if (ATTR) { at0 = o.a0; at1 = o.a1; }
        } else {
            const double c = (s == 3) ? 1.0 : 2.0; // s1 + 2 s2 + 2 s3 + s4, left to right (VK:959-960)
            hvel.x = hvel.x + c * o.hx;
            hvel.y = hvel.y + c * o.hy;
            hvel.z = hvel.z + c * o.hz;
            vvel = vvel + c * o.vv;
            if (ATTR) {
                at0 = at0 + c * o.a0;
                at1 = at1 + c * o.a1;
            }
        }
        hprev = mk3(o.hx, o.hy, o.hz);
    }
    io.status = st;
    io.hint_f = hint_f; io.hint_b = hint_b;
    if (st != ST_ALIVE) return io;
    if (P.use_euler) {
        new_pos = rotate_euler(pos, hvel, P.delta_t, r); // VK:968-972
    } else {
        { // (s1 + 2 s2 + 2 s3 + s4) / 6.0: six quotients by one constant
            const double a6[6] = {hvel.x, hvel.y, hvel.z, vvel, at0, at1};
            double q6[6];
            bool ok6 = true;
#pragma unroll
            for (int i = 0; i < 6; ++i) q6[i] = div_by6(a6[i], ok6);
            if (!ok6) {
#pragma unroll
                for (int i = 0; i < 6; ++i) q6[i] = slow_div(a6[i], 6.0);
            }
            hvel.x = q6[0]; hvel.y = q6[1]; hvel.z = q6[2];
            vvel = q6[3];
            at0 = q6[4]; at1 = q6[5];
        }
        const double tx = pos.x + hvel.x * dt, ty = pos.y + hvel.y * dt, tz = pos.z + hvel.z * dt; // VK:962-964
        const double tl = len3(tx, ty, tz);
        if (tl > 1e-12) {
            double ux, uy, uz;
            div3(tx, ty, tz, tl, ux, uy, uz);
            new_pos = mk3(ux * r, uy * r, uz * r);
        } else {
            new_pos = pos;
        }
    }
    // depth / radius update with the float round trip (VK:977-986, R3, R4)
    const double old_depth = (double)io.depth_f;
    double new_depth = old_depth - vvel * (double)P.delta_t;
    new_depth = (0.0 < new_depth) ? new_depth : 0.0;
    const double r_sum = r + vvel * (double)P.delta_t;
    const double r_new = (1.0 < r_sum) ? r_sum : 1.0;
    io.depth_f = (float)new_depth;
    const double nlen = len3(new_pos);
    if (nlen > 1e-12) {
        double ux, uy, uz;
        div3(new_pos.x, new_pos.y, new_pos.z, nlen, ux, uy, uz);
        new_pos = mk3(ux * r_new, uy * r_new, uz * r_new);
    }
    io.pos = new_pos;
    io.hvel = hvel;
    io.at0 = at0; io.at1 = at1;
    return io;
}

template <int M, bool PATH, bool EXTRA, bool ATTR>
__device__ __noinline__ StepIO step_generic_cold(const AdvectParams& P, int cell, int step, StepIO io)
{
    return step_generic<M, PATH, EXTRA, ATTR>(P, cell, step, io);
}

// EXTRA = false is the production instantiation; EXTRA = true additionally honours P.walk (MOPS_SEM_WALK) and P.diag_edge
// (near-edge counting) -- kept out of the hot variant because even never-taken branches cost registers and time.
// ATTR = true carries the pathline's scalar attributes (P.attr_count > 0 and an output buffer).
// SEG = true: the launch covers steps [P.step_begin, P.step_end) only (see AdvectParams::state); SEG = false is a
// single-launch kernel (every SEG-only branch folds away at compile time).
// NOW: how the vertical velocity of the call's snapshots is stored (0 present, 1 absent, 2 absent + zTop in the w slot), fastpath.cuh.
// FAST = true (hexagonal meshes, no attributes, no diagnostics): RK4 steps on hexagons with monotone columns run the
// straight-line form of fastpath.cuh; step_generic is called out of line for every step it does not cover.
template <int M, bool PATH, int MINB, bool EXTRA, bool ATTR, bool SEG = false, int NOW = 0, bool FAST = false>
__global__ void __launch_bounds__(MOPS_ADV_BLOCK, MINB) k_advect(const __grid_constant__ AdvectParams P)
{
    static_assert(!FAST || (M == 6 && !EXTRA && !ATTR), "the straight-line path is the hexagon / no-attribute / no-diagnostic form");
    const long long tix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long my_steps = 0, my_alive = 0, my_near = 0, my_above = 0;
    const long long n_act = (SEG && P.n_live) ? (long long)*P.n_live : P.n;
    if (tix < n_act) {
        const long long pid = P.order ? (long long)P.order[tix] : tix;
        const CellRec<M>* __restrict__ recs = reinterpret_cast<const CellRec<M>*>(P.rec);
        d3 pos = mk3(P.pos[3 * pid], P.pos[3 * pid + 1], P.pos[3 * pid + 2]);
        float depth_f = P.depth[pid];
        int cell = P.cell0[pid];
        const long long base = pid * (long long)P.each;
        int status = ST_ALIVE;
        int started = 0;
        int run_time = 0;
        int upd = 0;
        int hint_f = -1, hint_b = -1;
        bool first_vel = true;
        double edge_min = 1.0e300;
        const int step_lo = SEG ? P.step_begin : 0;
        const int step_hi = SEG ? P.step_end : P.times;
        const bool resume = SEG && step_lo > 0;
        if (resume) { // pick up where the previous segment parked this particle (it was alive and in a valid cell)
            const AdvState sv = P.state[pid];
            cell = sv.cell; upd = sv.upd; started = sv.started; hint_f = sv.hint_f; hint_b = sv.hint_b;
            first_vel = sv.first_vel != 0;
            run_time = step_lo * abs(P.delta_t);
        }
        const int started0 = started;
        // FAST: alpha = step / times and dalpha as exact quotients without nvcc's per-division branch
        const double x_times = FAST ? recip_refine((double)P.times) : 0.0;
        const double dalpha_f = (FAST && PATH) ? (double)P.delta_t / P.duration : 0.0; // VK:1401

        // Every output slot is written exactly once by this thread (no host-side memset of the buffers,
        // which may be pinned host memory written over PCIe): the reference's buffers are
        // value-initialised (TrajectoryCommon.h:20-25), so slots it never reaches hold (0,0,0).
        if (cell < 0 || cell >= P.nC) {
            status = ST_BAD_CELL; // VK:895-897: nothing is written, not even the seed
        } else {
            if (!resume) {
                st3(P.out_pos, base, pos.x, pos.y, pos.z); // VK:901
                st3(P.out_vel, base, 0.0, 0.0, 0.0);       // overwritten by the first completed step (VK:988-991)
                if (P.out_attr) st3(P.out_attr, base, 0.0, 0.0, 0.0);
            }
            for (int step = step_lo; step < step_hi; ++step) {
                run_time += abs(P.delta_t);
                if (step > 0) {
                    // relocation: argmin over {cellsOnCell[c][0..nv-1], c} of |centre - x|, strict <,
                    // that order, one ring (VK:903-921; GetCellNeighborsIdx TK:74-101)
                    const CellRec<M>* r = recs + cell;
                    double min_len = 1.7976931348623157e308;
                    int best = cell;
                    double len[M + 1]; // the candidates' distances: independent roots, one group
                    int cand[M + 1];
#pragma unroll
                    for (int k = 0; k < M; ++k) {
                        cand[k] = r->nbr[k]; // slots >= nv hold -1 (upload_records)
                        if (cand[k] >= 0) {
                            const double4 cc = ldg_d4(P.c4 + cand[k]);
                            const double dx = cc.x - pos.x, dy = cc.y - pos.y, dz = cc.z - pos.z;
                            len[k] = dx * dx + dy * dy + dz * dz;
                        } else {
                            len[k] = 1.0;
                        }
                    }
                    {
                        const double4 cc = ldg_d4(P.c4 + cell);
                        const double dx = cc.x - pos.x, dy = cc.y - pos.y, dz = cc.z - pos.z;
                        len[M] = dx * dx + dy * dy + dz * dz;
                        cand[M] = cell;
                    }
                    // argmin on the squared distances; sqrt is monotone, so the sqrt'd compare of the reference can
                    // only differ when two candidates round to the same root, i.e. lie within a few ulp of each
                    // other -- then (and only then) the roots are taken and compared as the reference does
                    double m1 = 1.7976931348623157e308;
#pragma unroll
                    for (int k = 0; k <= M; ++k)
                        if (cand[k] >= 0 && len[k] < m1) { m1 = len[k]; best = cand[k]; }
                    const double lim = m1 + m1 * 0x1p-48;
                    int close = 0;
#pragma unroll
                    for (int k = 0; k <= M; ++k) close += (cand[k] >= 0 && len[k] <= lim) ? 1 : 0;
                    if (close > 1 || !(m1 < 1.0e300)) {
                        best = cell;
                        sqrt_group<M + 1>(len);
#pragma unroll
                        for (int k = 0; k <= M; ++k)
                            if (cand[k] >= 0 && len[k] < min_len) { min_len = len[k]; best = cand[k]; }
                    }
                    if (best != cell) { cell = best; }
                    // walk mode: not limited to one ring (identical whenever the step is shorter than a cell)
                    if (EXTRA && P.walk) cell = walk_nearest<M>(recs, P.c4, cell, pos.x, pos.y, pos.z);
                }
                ++started;
                if (P.cell_log) P.cell_log[pid * (long long)P.times + step] = P.c_int2ext[cell];

                d3 hvel;
                double at0 = 0.0, at1 = 0.0;
                bool done = false;
                if (FAST) {
                    const CellRec<M>* __restrict__ rec = recs + cell;
                    bool covered = (rec->nv == M) & (P.sv[0].mono[cell] != 0) & (P.use_euler == 0);
                    if (PATH) covered = covered & (P.sv[1].mono[cell] != 0);
                    if (covered) {
                        const double alpha = PATH ? div_by((double)step, (double)P.times, x_times) : 0.0; // VK:1345 (exact: 1 <= times < 2^31)
                        FastStep fs;
                        const unsigned bad = fast_rk4_step<M, PATH, NOW>(rec, P.sv, P.L, pos, depth_f, alpha, dalpha_f, P.delta_t, hint_f, hint_b, fs);
                        if (bad == 0u) {
                            pos = fs.new_pos; hvel = fs.hvel; depth_f = fs.depth_f;
                            done = true;
                        }
                    }
                }
                if (!done) {
                    StepIO io;
                    io.pos = pos; io.hvel = mk3(0.0, 0.0, 0.0); io.at0 = 0.0; io.at1 = 0.0; io.edge_min = edge_min; io.depth_f = depth_f;
                    io.hint_f = hint_f; io.hint_b = hint_b; io.status = ST_ALIVE;
                    if (FAST) io = step_generic_cold<M, PATH, EXTRA, ATTR>(P, cell, step, io);
                    else io = step_generic<M, PATH, EXTRA, ATTR>(P, cell, step, io);
                    hint_f = io.hint_f; hint_b = io.hint_b;
                    if (EXTRA) edge_min = io.edge_min;
                    if (io.status != ST_ALIVE) { status = io.status; break; }
                    pos = io.pos; hvel = io.hvel; depth_f = io.depth_f;
                    if (ATTR) { at0 = io.at0; at1 = io.at1; }
                }

                if (first_vel) { // VK:988-991 / VK:1449-1456
                    first_vel = false;
                    st3(P.out_vel, base, hvel.x, hvel.y, hvel.z);
                    if (ATTR) st3(P.out_attr, base, at0, at1, 0.0);
                }
                bool rec_now;
                if (PATH) rec_now = (P.record_interval > 0) && (((step + 1) % P.record_interval) == 0); // VK:1470-1471
                else rec_now = (run_time % P.record_t) == 0;                                              // VK:994
                if (rec_now) {
                    if (upd < P.each) {
                        st3(P.out_pos, base + upd, pos.x, pos.y, pos.z);
                        st3(P.out_vel, base + upd, hvel.x, hvel.y, hvel.z);
                        if (ATTR) st3(P.out_attr, base + upd, at0, at1, 0.0);
                    }
                    ++upd;
                }
            }
        }
        P.pos[3 * pid] = pos.x; P.pos[3 * pid + 1] = pos.y; P.pos[3 * pid + 2] = pos.z;
        P.depth[pid] = depth_f;
        const bool park = SEG && status == ST_ALIVE && step_hi < P.times;
        if (park) { // alive with steps left: the next segment resumes from here
            AdvState sv;
            sv.cell = cell; sv.upd = upd; sv.started = started; sv.hint_f = hint_f; sv.hint_b = hint_b;
            sv.first_vel = first_vel ? 1 : 0; sv.alive = 1; sv.pad = 0;
            P.state[pid] = sv;
        } else {
            int k0 = (status == ST_BAD_CELL) ? 0 : ((upd < 1) ? 1 : ((upd < P.each) ? upd : P.each));
            for (int k = k0; k < P.each; ++k) {
                st3(P.out_pos, base + k, 0.0, 0.0, 0.0);
                st3(P.out_vel, base + k, 0.0, 0.0, 0.0);
                if (P.out_attr) st3(P.out_attr, base + k, 0.0, 0.0, 0.0);
            }
            if (SEG && P.state) P.state[pid].alive = 0;
            if (P.status) P.status[pid] = status;
            if (P.steps) P.steps[pid] = started;
            if (P.fcell) P.fcell[pid] = (cell >= 0 && cell < P.nC) ? P.c_int2ext[cell] : -1;
            if (EXTRA && P.min_edge) P.min_edge[pid] = edge_min;
            my_alive = (status == ST_ALIVE) ? 1ull : 0ull;
            my_near = (EXTRA && P.diag_edge && edge_min < 1e-12) ? 1ull : 0ull;
            my_above = (PATH && status == ST_ABOVE_SURFACE) ? 1ull : 0ull;
        }
        my_steps = (unsigned long long)(started - started0);
    }
    // one atomic pair per warp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        my_steps += __shfl_xor_sync(0xffffffffu, my_steps, off);
        my_alive += __shfl_xor_sync(0xffffffffu, my_alive, off);
        my_near += __shfl_xor_sync(0xffffffffu, my_near, off);
        if (PATH) my_above += __shfl_xor_sync(0xffffffffu, my_above, off);
    }
    if ((threadIdx.x & 31) == 0 && P.counters) {
        atomicAdd(P.counters + 0, my_steps);
        atomicAdd(P.counters + 1, my_alive);
        if (my_near) atomicAdd(P.counters + 3, my_near);
        if (PATH && my_above) atomicAdd(P.counters + 4, my_above);
    }
}

// =========================================================================================
// remap: one thread per pixel -- pixel -> lat/lon -> XYZ -> locate -> IsInMesh -> Wachspress
// -> depth test -> 2-layer velocity blend -> ENU (VisualizeFixedDepth, VK:238-471; replaces the
// serial host KD loop TBBKernel::SearchKDTree as well)
// =========================================================================================
struct RemapParams {
    const void* rec;
    const double4* c4;
    const int* cube;
    KdView kd;
    const int* c_int2ext;
    int F, nC, L;
    SnapView s;
    int attr_count;  // number of attribute arrays available (0..2)
    int attr_image;  // 1 = write img1 (reference: mDoubleAttributes.size() > 1)
    int width, height;
    double minLat, maxLat, minLon, maxLon;
    double DEPTH;    // -FixedDepth
    double* img0;
    double* img1;
    int* pixel_cell; // caller ids or null
    unsigned long long* nan_count;
};

__device__ __forceinline__ void put_pixel(double* img, long long gid, double a, double b, double c)
{
    double4* p = reinterpret_cast<double4*>(img) + gid; // (i*w + j)*4 doubles; cudaMalloc'd => 32 B aligned
    *p = make_double4(a, b, c, 1.0);
}

template <int M>
__global__ void __launch_bounds__(128) k_remap(const RemapParams P)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.width * P.height) return;
    const int ih = (int)(gid / P.width);
    const int jw = (int)(gid % P.width);
    // GeoConverter::convertPixelToLatLonToRadians + convertRadianLatLonToXYZ (GeoConverter.hpp:9-33,107-125)
    double lat = P.maxLat - ((double)ih / (double)P.height * (P.maxLat - P.minLat));
    double lon = ((double)jw / (double)P.width * (P.maxLon - P.minLon)) + P.minLon;
    lat = lat * (3.14159265358979323846 / 180.0);
    lon = lon * (3.14159265358979323846 / 180.0);
    const double rr = 6371010.0;
    double sintheta, costheta, sinphi, cosphi;
    sincos(lat, &sintheta, &costheta);
    sincos(lon, &sinphi, &cosphi);
    d3 pos;
    pos.x = rr * costheta * cosphi;
    pos.y = rr * costheta * sinphi;
    pos.z = rr * sintheta;

    const CellRec<M>* __restrict__ recs = reinterpret_cast<const CellRec<M>*>(P.rec);
    const double nanv = nan("");
    int cell = -1;
    if (finite3(pos.x, pos.y, pos.z)) cell = locate_cell<M>(recs, P.c4, P.cube, P.F, P.kd, pos.x, pos.y, pos.z);
    if (P.pixel_cell) P.pixel_cell[gid] = cell >= 0 ? P.c_int2ext[cell] : -1;

    bool ok = (cell >= 0 && cell < P.nC);
    double u_east = 0.0, v_north = 0.0, spd = 0.0, a0 = 0.0, a1 = 0.0;
    if (ok) {
        const CellRec<M>* __restrict__ rec = recs + cell;
        const int nv = rec->nv;
        const int L = P.L;
        double w[M];
        bool wfinite = false;
        ok = (nv > 0) && cell_weights<M>(rec, nv, pos.x, pos.y, pos.z, w, wfinite);
        if (ok) {
            voff_t vo[M];
#pragma unroll
            for (int i = 0; i < M; ++i) vo[i] = (i < nv) ? (voff_t)rec->vid[i] * (voff_t)L : 0u;
            const double DEPTH = P.DEPTH;
            int local_layer = -1;
            double topI = 0.0, botI = 0.0;
            if ((P.s.mono[cell] != 0) && wfinite) {
                // non-increasing column: only z[0], z[1], z[L-1] are ever needed
                const ZCol<M> z{P.s.ztop, w, vo, nv, L};
                const double zs = z(0), zb = z(L - 1);
                double z_surf = zs, z_bot = zb;
                if (z_surf < z_bot) { const double t = z_surf; z_surf = z_bot; z_bot = t; }
                const double ad = 1e-8 * fabs(z_surf - z_bot);
                const double epsd = (1e-6 < ad) ? ad : 1e-6;
                if (!(DEPTH <= z_surf + epsd && DEPTH >= z_bot - epsd)) ok = false;
                if (ok) {
                    if (DEPTH <= zs) {
                        local_layer = 0; // VK:392-394
                        topI = zs; botI = zs;
                    } else {
                        const double z1 = z(1);
                        // DEPTH > z[0] >= z[k]: layer 1 matches iff DEPTH <= z[0]+1e-8; if it does not,
                        // no later layer can (their tops are <= z[0]) and the pixel is NaN (VK:378-401)
                        if (DEPTH <= zs + 1e-8 && DEPTH >= z1 - 1e-8) { local_layer = 1; topI = zs; botI = z1; }
                    }
                }
            } else {
                const LayerRes lr = slow_layer_remap<M>(rec, P.s.ztop, L, pos.x, pos.y, pos.z, DEPTH);
                local_layer = lr.layer; topI = lr.top; botI = lr.bot;
                if (lr.layer == -2) { ok = false; local_layer = -1; }
            }
            if (local_layer < 0) ok = false;
            if (ok) {
                if (topI < botI) { const double t = topI; topI = botI; botI = t; }
                const double denom = topI - botI;
                const double tparam = (denom > 1e-12) ? (DEPTH - botI) / denom : 0.5; // VK:411-412
                int j = local_layer - 1;
                if (j < 0) j = 0;
                if (j > L - 1) j = L - 1;
                const int j_bot = (j + 1 < L - 1) ? j + 1 : L - 1;
                const int j_top = j;
                double tx, ty, tz, tw, bx, by, bz, bw;
                gather_velw<M>(P.s.velw, vo, w, nv, j_top, tx, ty, tz, tw);
                gather_velw<M>(P.s.velw, vo, w, nv, j_bot, bx, by, bz, bw);
                const double mtop = len3(tx, ty, tz), mbot = len3(bx, by, bz);
                double fx, fy, fz;
                if (mtop < 1e-12 && mbot < 1e-12) { fx = 0.0; fy = 0.0; fz = 0.0; }
                else if (mtop < 1e-12) { fx = bx; fy = by; fz = bz; }
                else if (mbot < 1e-12) { fx = tx; fy = ty; fz = tz; }
                else {
                    const double omt = 1.0 - tparam;
                    fx = omt * bx + tparam * tx; // VK:435
                    fy = omt * by + tparam * ty;
                    fz = omt * bz + tparam * tz;
                }
                // GeoConverter::convertXYZVelocityToENU (GeoConverter.hpp:200-223)
                if (pos.x == 0.0 && pos.y == 0.0) {
                    u_east = 0.0; v_north = 0.0;
                } else {
                    const double Rxy = sqrt(pos.x * pos.x + pos.y * pos.y);
                    const double Rxyz = sqrt(pos.x * pos.x + pos.y * pos.y + pos.z * pos.z);
                    const double slon = pos.y / Rxy, clon = pos.x / Rxy, slat = pos.z / Rxyz, clat = Rxy / Rxyz;
                    u_east = -slon * fx + clon * fy;
                    v_north = -slat * (clon * fx + slon * fy) + clat * fz;
                }
                spd = sqrt(u_east * u_east + v_north * v_north);
                if (P.attr_image) { // VK:443-464: attributes from layer max(local_layer-1, 0)
                    if (P.attr_count >= 1) a0 = gather_scalar<M>(P.s.attr0, vo, w, nv, j);
                    if (P.attr_count >= 2) a1 = gather_scalar<M>(P.s.attr1, vo, w, nv, j);
                }
            }
        }
    }
    if (ok) {
        put_pixel(P.img0, gid, u_east, v_north, spd);
        if (P.attr_image && P.img1) put_pixel(P.img1, gid, a0, a1, 0.0);
    } else {
        put_pixel(P.img0, gid, nanv, nanv, nanv);
        if (P.attr_image && P.img1) put_pixel(P.img1, gid, nanv, nanv, nanv);
        if (P.nan_count) atomicAdd(P.nan_count, 1ull);
    }
}

// =========================================================================================
// the two other lat/lon views of the reference, same locate + Wachspress + gather
// building blocks, one thread per pixel:
//   MODE 0  VisualizeFixedLayer    VK:141-236   velocity of ONE layer -> ENU, pixel (u_east, v_north, 0, 1)
//   MODE 1  VisualizeFixedLatitude VK:473-651   depth-vs-longitude section at a fixed latitude
//           (host loops only in the reference's CUDA backend)
// =========================================================================================
struct ViewParams {
    const void* rec;
    const double4* c4;
    const int* cube;
    KdView kd;
    const int* c_int2ext;
    int F, nC, L;
    SnapView s;
    int width, height;
    double minLat, maxLat, minLon, maxLon; // MODE 0 pixel -> lat/lon as in k_remap
    int fixed_layer;                       // MODE 0 (already clamped to [0, L-1])
    double fixed_lat, minDepth, maxDepth;  // MODE 1: refBottomDepth front / back
    double* img;
    int* pixel_cell;
    unsigned long long* nan_count;
};

template <int M, int MODE>
__global__ void __launch_bounds__(128) k_view(const ViewParams P)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.width * P.height) return;
    const int ih = (int)(gid / P.width);
    const int jw = (int)(gid % P.width);
    const double rr = 6371010.0;
    double lat, lon, DEPTH = 0.0;
    if (MODE == 0) {
        lat = P.maxLat - ((double)ih / (double)P.height * (P.maxLat - P.minLat));
        lon = ((double)jw / (double)P.width * (P.maxLon - P.minLon)) + P.minLon;
        lat = lat * (3.14159265358979323846 / 180.0);
        lon = lon * (3.14159265358979323846 / 180.0);
    } else {
        const double i_step = (P.height > 1) ? (P.maxDepth - P.minDepth) / (P.height - 1) : 0.0; // VK:513-514
        const double j_step = (P.width > 1) ? (P.maxLon - P.minLon) / (P.width - 1) : 0.0;
        const double depth_plot = P.minDepth + ih * i_step;
        DEPTH = -fabs(depth_plot);
        lat = P.fixed_lat * (3.14159265358979323846 / 180.0);
        lon = (P.minLon + jw * j_step) * (3.14159265358979323846 / 180.0);
    }
    double sintheta, costheta, sinphi, cosphi;
    sincos(lat, &sintheta, &costheta);
    sincos(lon, &sinphi, &cosphi);
    d3 pos;
    pos.x = rr * costheta * cosphi;
    pos.y = rr * costheta * sinphi;
    pos.z = rr * sintheta;

    const CellRec<M>* __restrict__ recs = reinterpret_cast<const CellRec<M>*>(P.rec);
    int cell = -1;
    if (finite3(pos.x, pos.y, pos.z)) cell = locate_cell<M>(recs, P.c4, P.cube, P.F, P.kd, pos.x, pos.y, pos.z);
    if (P.pixel_cell) P.pixel_cell[gid] = cell >= 0 ? P.c_int2ext[cell] : -1;

    bool ok = (cell >= 0 && cell < P.nC);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (ok) {
        const CellRec<M>* __restrict__ rec = recs + cell;
        const int nv = rec->nv;
        const int L = P.L;
        ok = nv > 0;
        if (ok) {
            if (MODE == 0) {
                ok = in_mesh<M>(rec, nv, pos.x, pos.y, pos.z);
            } else {
                // MPASOField::isOnOcean (src/Core/MPASOField.cpp:36-81): dot(cross(A,B), p - A) must have
                // one sign for every edge (cross(O-A, O-B) == cross(A,B) bit for bit)
                bool first_pos = false, land = false;
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    if (k < nv) {
                        const double dir = rec->nx[k] * (pos.x - rec->vx[k]) + rec->ny[k] * (pos.y - rec->vy[k]) + rec->nz[k] * (pos.z - rec->vz[k]);
                        const bool positive = dir > 0;
                        if (k == 0) first_pos = positive;
                        else if (positive != first_pos) land = true;
                    }
                }
                ok = !land;
            }
        }
        if (ok) {
            double w[M];
            bool wfinite = false;
            wachspress_weights<M>(rec, nv, pos.x, pos.y, pos.z, w, wfinite);
            voff_t vo[M];
#pragma unroll
            for (int i = 0; i < M; ++i) vo[i] = (i < nv) ? (voff_t)rec->vid[i] * (voff_t)L : 0u;
            double dummy;
            if (MODE == 0) {
                gather_velw<M>(P.s.velw, vo, w, nv, P.fixed_layer, fx, fy, fz, dummy); // VK:220-226
            } else {
                // column + fix-up (VK:566-582), range test and first-match scan with EPSILON = 1e-6 (VK:584-604)
                const double EPS = 1e-6;
                int layer = -1;
                double z_up = 0.0, z_dn = 0.0;
                if ((P.s.mono[cell] != 0) && wfinite) {
                    const ZCol<M> z{P.s.ztop, w, vo, nv, L};
                    if (DEPTH > z(0) + EPS || DEPTH < z(L - 1) - EPS) ok = false;
                    if (ok) {
                        // non-increasing column: first k with DEPTH >= z[k] - EPS (DEPTH <= z[k-1] + EPS then holds)
                        int lo = 1, hi = L - 1;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (DEPTH >= z(mid) - EPS) hi = mid;
                            else lo = mid + 1;
                        }
                        layer = lo;
                        z_up = z(layer - 1);
                        z_dn = z(layer);
                    }
                } else {
                    const LayerRes lr = slow_layer_latitude<M>(rec, P.s.ztop, L, pos.x, pos.y, pos.z, DEPTH);
                    layer = lr.layer; z_up = lr.top; z_dn = lr.bot;
                    if (layer < 0) ok = false;
                }
                if (ok) {
                    if (z_up < z_dn) { const double t = z_up; z_up = z_dn; z_dn = t; }
                    const double denom = z_up - z_dn;
                    if (fabs(denom) < 1e-30) ok = false; // VK:616-620
                    if (ok) {
                        const double t = (DEPTH - z_dn) / denom; // not clamped (VK:622)
                        double ux, uy, uz, dx, dy, dz;
                        gather_velw<M>(P.s.velw, vo, w, nv, layer - 1, ux, uy, uz, dummy);
                        gather_velw<M>(P.s.velw, vo, w, nv, layer, dx, dy, dz, dummy);
                        const double omt = 1.0 - t;
                        fx = omt * dx + t * ux; // VK:641
                        fy = omt * dy + t * uy;
                        fz = omt * dz + t * uz;
                    }
                }
            }
        }
    }
    if (ok) {
        double u_east = 0.0, v_north = 0.0;
        if (!(pos.x == 0.0 && pos.y == 0.0)) { // GeoConverter::convertXYZVelocityToENU
            const double Rxy = sqrt(pos.x * pos.x + pos.y * pos.y);
            const double Rxyz = sqrt(pos.x * pos.x + pos.y * pos.y + pos.z * pos.z);
            const double slon = pos.y / Rxy, clon = pos.x / Rxy, slat = pos.z / Rxyz, clat = Rxy / Rxyz;
            u_east = -slon * fx + clon * fy;
            v_north = -slat * (clon * fx + slon * fy) + clat * fz;
        }
        put_pixel(P.img, gid, u_east, v_north, 0.0);
    } else {
        const double nanv = nan("");
        put_pixel(P.img, gid, nanv, nanv, nanv);
        if (P.nan_count) atomicAdd(P.nan_count, 1ull);
    }
}

} // namespace mops
