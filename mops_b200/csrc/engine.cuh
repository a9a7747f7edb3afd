// engine.cuh -- device-side data layout and the per-point evaluation shared by the
// streamline / pathline / remap kernels.
//
// Reference being re-designed (paths relative to the reference root; VK =
// src/CPU/TBB/Kernel/MPASOVisualizerKernels.cpp, TK = src/CPU/TBB/Kernel/TBBKernel.h):
//   calc_velocity_at (streamline) VK:740-872, (pathline) VK:1124-1327, remap pixel VK:288-470,
//   IsInMesh TK:21-54, CalcPolygonWachspress src/Utils/Interpolation.hpp:137-165,
//   CalcVelocity / CalcAttribute TK:128-164.
//
// What is different from the reference (and why results are still identical):
//  * one packed, 32-byte-aligned record per cell (CellRec) holds everything a particle
//    needs from the mesh: vertex ids, neighbour ids, vertex positions, and two
//    point-independent products the reference recomputes for every evaluation -- the edge
//    normals cross(v_k, v_k+1) of IsInMesh and the Wachspress corner areas B_i.  They are
//    computed once per mesh by the same expressions, so every later result is bit-equal.
//  * the depth column is never materialised (the reference keeps double[100] per thread):
//    when every vertex column of the cell is non-increasing (checked once per snapshot)
//    the interpolated column is non-increasing too (weights are >= 0, rounding is
//    monotone), the reference's fix-up `z[k] = z[k-1] - 1e-9` is a no-op, and the layer
//    search may probe levels on the fly.  A layer hint from the previous evaluation is
//    accepted only when it is provably the unique answer of the reference's search;
//    otherwise the reference's search runs.  Cells that fail the check take a streaming
//    restatement of the full-column path.
//  * velocity (x,y,z) and vertical velocity share one 32-byte (vertex, level) record.
#pragma once
#include "dmath.cuh"
#include <stdint.h>

namespace mops {

// row offset (vertex id * L) into the vertex-major snapshot arrays: unsigned, so address arithmetic needs no sign word
typedef unsigned voff_t;

// ---- resident mesh ----------------------------------------------------------------------
template <int M>
struct alignas(32) CellRec {
    int nv;       // nEdgesOnCell; 0 when the cell is unusable (nv > M cannot happen by construction)
    int pad;
    int vid[M];   // internal vertex ids (0-based), -1 beyond nv
    int nbr[M];   // internal neighbour cell ids in cellsOnCell order, -1 = none
    double vx[M], vy[M], vz[M]; // vertex positions
    double nx[M], ny[M], nz[M]; // cross(v_k, v_(k+1)%nv)                      (TK:45)
    double B[M];                // triangle_area(v_(i-1), v_i, v_(i+1))       (Interpolation.hpp:154)
};

struct VertRec {      // per Voronoi vertex, mesh-constant part of the cell->vertex interpolation
    int c0, c1, c2;   // internal ids of cellsOnVertex
    int boundary;     // reference's `(id-1) > nCells+1` test (MPASOSolutionTBB.cpp:35)
    double u, v, w;   // calcTriangleBarycentric of the vertex in (c0,c1,c2) (Interpolation.hpp:79-93)
};

struct SnapView {
    const double* __restrict__ ztop;        // [nV][L]   cellVertexZTop, internal vertex order
    const double4* __restrict__ velw;       // [nV][L]   (vx,vy,vz, vertVelocityTop level k)
    const double* __restrict__ attr0;       // [nV][L] or null
    const double* __restrict__ attr1;
    const unsigned char* __restrict__ mono; // [nC] 1 = all vertex columns of the cell non-increasing
    int w_is_z; // 1: the snapshot was uploaded without vertVelocityTop and the w slot of every velw record holds zTop of the
                // same (vertex, level) instead of +0.0 (one 32-byte load then serves velocity AND the layer probe of the
                // straight-line path); every reader of w substitutes 0.0
};

enum {
    ST_ALIVE = 0, ST_BAD_CELL = 1, ST_NOT_IN_CELL = 2, ST_BAD_COLUMN = 3, ST_ZERO_VELOCITY = 4,
    ST_ABOVE_SURFACE = 5, ST_BAD_SETUP = 6,
    ST_GENERIC = -1 // internal: the hexagon fast path does not cover this evaluation, take the generic one
};

// ---- Wachspress weights + in-cell test ------------------------------------------------
// returns false when the point is not in the cell (IsInMesh).  w[] are the normalised
// weights, vo[i] = vid[i]*L (row offsets into the vertex-major arrays); wfinite tells
// whether all weights are finite (a point exactly on an edge gives inf/NaN, Appendix B N7).
// Diagnostic only (cfg.count_near_edge): smallest angular distance of p to the great circles through
// the cell's edges, |dot(n_k, p)| / (|n_k| |p|) ~ angle for small angles.
template <int M>
__device__ __noinline__ double min_edge_angle(const CellRec<M>* __restrict__ rec, int nv, double px, double py, double pz)
{
    double best = 1.0e300;
    const double pl = len3(px, py, pz);
    for (int k = 0; k < nv; ++k) {
        const double direction = rec->nx[k] * px + rec->ny[k] * py + rec->nz[k] * pz;
        const double a = fabs(direction) / (len3(rec->nx[k], rec->ny[k], rec->nz[k]) * pl);
        if (a < best) best = a;
    }
    return best;
}

// TBBKernel::IsInMesh on the precomputed edge normals (TK:21-54)
// FULL = true: the caller knows nv == M (every slot of the record is a real vertex), so the per-slot
// `k < nv` selects and the wrap-around selects fold away at compile time -- the hexagon fast path.
template <int M, bool FULL = false>
__device__ __forceinline__ bool in_mesh(const CellRec<M>* __restrict__ rec, int nv_in, double px, double py, double pz)
{
    const int nv = FULL ? M : nv_in;
    if (!finite3(px, py, pz)) return false; // TK:29-33
    bool inside = true;
#pragma unroll
    for (int k = 0; k < M; ++k) {
        if (k < nv) {
            const double direction = rec->nx[k] * px + rec->ny[k] * py + rec->nz[k] * pz; // TK:46
            if (direction < 0.0) inside = false;
        }
    }
    return inside;
}

// Interpolator::CalcPolygonWachspress (src/Utils/Interpolation.hpp:137-165) with the corner areas B_i
// taken from the record.  wfinite: all weights finite (a point exactly on an edge gives inf/NaN, N7).
template <int M, bool FULL = false>
__device__ __forceinline__ void wachspress_weights(const CellRec<M>* __restrict__ rec, int nv_in, double px, double py, double pz,
                                                   double (&w)[M], bool& wfinite)
{
    const int nv = FULL ? M : nv_in;
    double ax[M], ay[M], az[M];
#pragma unroll
    for (int k = 0; k < M; ++k) {
        ax[k] = rec->vx[k]; ay[k] = rec->vy[k]; az[k] = rec->vz[k];
    }
    double ar[M]; // ar[k] = area(v_k, v_(k+1)%nv, p); the M roots are independent -> one group
#pragma unroll
    for (int k = 0; k < M; ++k) {
        if (k < nv) {
            const bool wrap = (k + 1 >= nv);
            const double bx = wrap ? ax[0] : ax[(k + 1) % M];
            const double by = wrap ? ay[0] : ay[(k + 1) % M];
            const double bz = wrap ? az[0] : az[(k + 1) % M];
            ar[k] = tri_cross2(ax[k], ay[k], az[k], bx, by, bz, px, py, pz);
        } else {
            ar[k] = 1.0;
        }
    }
    sqrt_group<M>(ar);
#pragma unroll
    for (int k = 0; k < M; ++k) ar[k] = (k < nv) ? ar[k] * 0.5 : 0.0; // '/ 2.0' of triangle_area (exact either way)
    double prev = ar[0]; // A_i of i = 0 is area(v_(nv-1), v_0, p) = ar[nv-1]
#pragma unroll
    for (int k = 1; k < M; ++k)
        if (k == nv - 1) prev = ar[k];
    double den[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        if (i < nv) {
            w[i] = rec->B[i];
            den[i] = prev * ar[i];
            prev = ar[i];
        } else {
            w[i] = 0.0;
            den[i] = 1.0;
        }
    }
    div_group<M>(w, den); // w_i = B_i / (A_i * A_(i+1)), independent quotients
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i)
        if (i < nv) sum += w[i];
    const double recp = 1.0 / sum;
#pragma unroll
    for (int i = 0; i < M; ++i)
        if (i < nv) w[i] *= recp;
    wfinite = isfinite(sum) && isfinite(recp) && (sum > 0.0);
}

// Hexagon fast path of CalcPolygonWachspress (nv == M): the same quotients with two point-independent pieces removed.
//    (keeping the edge v_(k+1) - v_k of triangle_area in the record was measured slower, profiles/r02_ab_fast2.txt: the L1
//    data pipe is the kernel's second roof and 144 more loaded bytes per evaluation cost more than 18 subtractions);
//  * the '/ 2.0' of the six triangle areas is dropped: scaling by a power of two commutes with rounding, so with
//    a_k = 2 A_k every quotient is fl(B_i / (a_(i-1) a_i)) = w_i / 4 exactly, the sum is sum / 4, its reciprocal 4 / sum and
//    the normalised weights (w_i / 4) * (4 / sum) are bit-identical to the reference's -- as long as nothing under- or
//    overflows, which the range windows below guarantee (a group outside them is not covered: ok = false);
//  * roots, quotients and the reciprocal run as the branch-free exact sequences of dmath.cuh with ONE range decision at the
//    end, so there is no slow-path merge inside the hot code.
// ok = false: not covered (operands outside the windows, e.g. a point exactly on an edge) -> caller takes the generic path.
// LDG = true: the record is read with explicit ld.global.nc (the straight-line path launders `rec`, after which plain
// dereferences compile to generic-space loads).
template <int M, bool LDG = false>
__device__ __forceinline__ void hex_weights(const CellRec<M>* __restrict__ rec, double px, double py, double pz, double (&w)[M], bool& ok)
{
    auto ld = [](const double* p) { return LDG ? ldg_f64(p) : *p; };
    double a[M];
    {
        double vx[M], vy[M], vz[M];
#pragma unroll
        for (int k = 0; k < M; ++k) { vx[k] = ld(&rec->vx[k]); vy[k] = ld(&rec->vy[k]); vz[k] = ld(&rec->vz[k]); }
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const int kn = (k + 1) % M;
            a[k] = tri_cross2(vx[k], vy[k], vz[k], vx[kn], vy[kn], vz[kn], px, py, pz);
        }
    }
    unsigned mn = hi_raw(a[0]), mx = mn;
#pragma unroll
    for (int k = 1; k < M; ++k) {
        mn = min(mn, hi_raw(a[k]));
        mx = max(mx, hi_raw(a[k]));
    }
    ok = (mn >= 0x03500000u && mx < 0x7ff00000u); // nvcc's own fast-path range of sqrt
#pragma unroll
    for (int k = 0; k < M; ++k) a[k] = sq_fast(a[k]);
    double den[M];
#pragma unroll
    for (int i = 0; i < M; ++i) den[i] = a[(i + M - 1) % M] * a[i]; // A_i of vertex i is area(v_(i-1), v_i, p)
    unsigned wmn = 0xffffffffu, wmx = 0u;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        w[i] = div_by(ld(&rec->B[i]), den[i], recip_refine(den[i]));
        wmn = min(wmn, min(hi_raw(den[i]), hi_raw(w[i])));
        wmx = max(wmx, max(hi_raw(den[i]), hi_raw(w[i])));
    }
    double sum = w[0]; // reference: sum = 0.0; sum += w[0] -> 0.0 + w[0] == w[0] for w[0] > 0
#pragma unroll
    for (int i = 1; i < M; ++i) sum += w[i];
    const double recp = div_by(1.0, sum, recip_refine(sum));
    wmn = min(wmn, min(hi_raw(sum), hi_raw(recp)));
    wmx = max(wmx, max(hi_raw(sum), hi_raw(recp)));
    ok = ok && (wmn >= WIN_LO && wmx < WIN_HI);
#pragma unroll
    for (int i = 0; i < M; ++i) w[i] *= recp;
}

// returns false when the point is not in the cell (IsInMesh); else fills the normalised weights
template <int M, bool FULL = false>
__device__ __forceinline__ bool cell_weights(const CellRec<M>* __restrict__ rec, int nv, double px, double py, double pz,
                                             double (&w)[M], bool& wfinite)
{
    if (!in_mesh<M, FULL>(rec, nv, px, py, pz)) return false;
    wachspress_weights<M, FULL>(rec, nv, px, py, pz, w, wfinite);
    return true;
}

// ---- interpolated zTop column ---------------------------------------------------------------
struct LayerRes {
    int layer;       // reference's local_layer; for the pathline 0 = above surface, -1 = none
    double top, bot; // z[layer-1], z[layer] (after the fix-up)
};

// FAST PATH: the column is provably non-increasing, so raw value == fixed-up value and levels
// are probed on the fly with the weights / row offsets held in registers.
template <int M, bool FULL = false>
struct ZCol {
    const double* __restrict__ ztop;
    const double (&w)[M];
    const voff_t (&vo)[M];
    int nv;
    int L;
    __device__ __forceinline__ double operator()(int k) const
    {
        double z = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i)
            if (FULL || i < nv) z += w[i] * ztop[vo[i] + k]; // VK:774-781, accumulated in vertex order
        return z;
    }
    // z(k-1) and z(k) in one pass: one row pointer per vertex, the upper level at a fixed -8 B offset (each
    // accumulator still sums in vertex order, so both values are bit-equal to operator())
    __device__ __forceinline__ void pair(int k, double& top, double& bot) const
    {
        top = 0.0; bot = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) {
            if (FULL || i < nv) {
                const double* __restrict__ q = ztop + (vo[i] + (voff_t)k);
                top += w[i] * q[-1];
                bot += w[i] * q[0];
            }
        }
    }
};

// SLOW PATH (cells with a non-monotone vertex column, or non-finite weights): the reference's
// full column with its prefix-dependent fix-up (VK:772-789), materialised in a local array
// inside NON-INLINED functions that recompute the weights themselves (same expressions, so
// bit-identical) -- nothing is handed over through local memory, and the fast path keeps its
// arrays in registers.
template <int M>
__device__ __forceinline__ void fill_fixed_column(const CellRec<M>* __restrict__ rec, const double* __restrict__ ztop, int L,
                                                  double px, double py, double pz, double* col)
{
    double w[M];
    bool wfinite;
    const int nv = rec->nv;
    cell_weights<M>(rec, nv, px, py, pz, w, wfinite);
    for (int k = 0; k < L; ++k) {
        double z = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i)
            if (i < nv) z += w[i] * ztop[rec->vid[i] * L + k];
        col[k] = z;
    }
    for (int k = 1; k < L; ++k)
        if (col[k] > col[k - 1]) col[k] = col[k - 1] - 1e-9;
}

// streamline layer search, VK:791-822 (binary, eps = 1e-8) on any column accessor
template <class Z>
__device__ __forceinline__ LayerRes binary_layer_search(const Z& z, int L, double d)
{
    const double eps = 1e-8;
    int layer;
    if (d > z(0) + eps) {
        layer = 1;
    } else if (d < z(L - 1) - eps) {
        layer = L - 1;
    } else {
        int lo = 1, hi = L - 1, ans = 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const double top_i = z(mid - 1);
            const double bot_i = z(mid);
            if (d <= top_i + eps && d >= bot_i - eps) {
                ans = mid;
                break;
            }
            if (d > top_i + eps) hi = mid - 1;
            else lo = mid + 1;
        }
        if (ans < 1) ans = 1;
        if (ans > L - 1) ans = L - 1;
        layer = ans;
    }
    LayerRes r;
    r.layer = layer;
    z.pair(layer, r.top, r.bot);
    return r;
}

struct ArrayCol {
    const double* col;
    __device__ __forceinline__ double operator()(int k) const { return col[k]; }
    __device__ __forceinline__ void pair(int k, double& top, double& bot) const { top = col[k - 1]; bot = col[k]; }
};

template <int M>
__device__ __noinline__ LayerRes slow_layer_stream(const CellRec<M>* __restrict__ rec, const double* __restrict__ ztop, int L,
                                                   double px, double py, double pz, double d)
{
    double col[100];
    fill_fixed_column<M>(rec, ztop, L, px, py, pz, col);
    return binary_layer_search(ArrayCol{col}, L, d);
}

// `hint` (a previous answer, or < 1) is accepted only if it is the unique matching layer, in
// which case the reference's search returns it as well (proof in DESIGN.md, "layer hint").
template <int M, bool FULL>
__device__ __forceinline__ LayerRes layer_search_stream(const ZCol<M, FULL>& z, double d, int hint)
{
    const double eps = 1e-8;
    const int L = z.L;
    if (hint >= 1 && hint <= L - 1) {
        LayerRes r;
        r.layer = hint;
        z.pair(hint, r.top, r.bot);
        if (d <= r.top + eps && d >= r.bot - eps && d > r.bot + eps && d < r.top - eps) return r;
    }
    return binary_layer_search(z, L, d);
}

// pathline layer search, VK:1182-1218: above surface -> 0 (caller reports ABOVE_SURFACE),
// below bottom -> L-1, else the FIRST k in 1..L-1 with d <= z[k-1]+eps && d >= z[k]-eps.
template <int M>
__device__ __noinline__ LayerRes slow_layer_path(const CellRec<M>* __restrict__ rec, const double* __restrict__ ztop, int L,
                                                 double px, double py, double pz, double d)
{
    double col[100];
    fill_fixed_column<M>(rec, ztop, L, px, py, pz, col);
    const double eps = 1e-8;
    LayerRes r;
    r.layer = -1; r.top = 0.0; r.bot = 0.0;
    if (d > col[0] + eps) { r.layer = 0; return r; }
    if (d < col[L - 1] - eps) {
        r.layer = L - 1;
    } else {
        for (int k = 1; k < L; ++k)
            if (d <= col[k - 1] + eps && d >= col[k] - eps) { r.layer = k; break; }
    }
    if (r.layer >= 1) { r.top = col[r.layer - 1]; r.bot = col[r.layer]; }
    return r;
}

// remap (VisualizeFixedDepth) column logic on the fixed-up column, VK:346-409.  layer = -2: depth
// outside [z_bot - epsd, z_surf + epsd]; -1: no layer; else local_layer with top = z[max(0,l-1)].
template <int M>
__device__ __noinline__ LayerRes slow_layer_remap(const CellRec<M>* __restrict__ rec, const double* __restrict__ ztop, int L,
                                                  double px, double py, double pz, double DEPTH)
{
    double col[100];
    fill_fixed_column<M>(rec, ztop, L, px, py, pz, col);
    LayerRes r;
    r.layer = -1; r.top = 0.0; r.bot = 0.0;
    double z_surf = col[0], z_bot = col[L - 1];
    if (z_surf < z_bot) { const double t = z_surf; z_surf = z_bot; z_bot = t; }
    const double ad = 1e-8 * fabs(z_surf - z_bot);
    const double epsd = (1e-6 < ad) ? ad : 1e-6;
    if (!(DEPTH <= z_surf + epsd && DEPTH >= z_bot - epsd)) { r.layer = -2; return r; }
    for (int k = 1; k < L; ++k) {
        double tI = col[k - 1], bI = col[k];
        if (tI < bI) { const double t = tI; tI = bI; bI = t; }
        if (DEPTH <= tI + 1e-8 && DEPTH >= bI - 1e-8) { r.layer = k; break; }
    }
    if (DEPTH <= col[0]) r.layer = 0;
    if (r.layer < 0) return r;
    r.top = col[(r.layer - 1 > 0) ? r.layer - 1 : 0];
    r.bot = col[r.layer];
    return r;
}

// fixed-latitude section (VisualizeFixedLatitude) column logic on the fixed-up column, VK:566-604:
// range test and first-match scan with EPSILON = 1e-6; layer = -1: outside / none.
template <int M>
__device__ __noinline__ LayerRes slow_layer_latitude(const CellRec<M>* __restrict__ rec, const double* __restrict__ ztop, int L,
                                                     double px, double py, double pz, double DEPTH)
{
    double col[100];
    // the reference computes these weights WITHOUT the IsInMesh gate (it used isOnOcean before)
    {
        double w[M];
        bool wfinite;
        const int nv = rec->nv;
        wachspress_weights<M>(rec, nv, px, py, pz, w, wfinite);
        for (int k = 0; k < L; ++k) {
            double z = 0.0;
#pragma unroll
            for (int i = 0; i < M; ++i)
                if (i < nv) z += w[i] * ztop[rec->vid[i] * L + k];
            col[k] = z;
        }
        for (int k = 1; k < L; ++k)
            if (col[k] > col[k - 1]) col[k] = col[k - 1] - 1e-9;
    }
    const double EPS = 1e-6;
    LayerRes r;
    r.layer = -1; r.top = 0.0; r.bot = 0.0;
    if (DEPTH > col[0] + EPS || DEPTH < col[L - 1] - EPS) return r;
    for (int k = 1; k < L; ++k) {
        double zu = col[k - 1], zd = col[k];
        if (zu < zd) { const double t = zu; zu = zd; zd = t; }
        if (DEPTH <= zu + EPS && DEPTH >= zd - EPS) { r.layer = k; break; }
    }
    if (r.layer < 0) return r;
    r.top = col[r.layer - 1];
    r.bot = col[r.layer];
    return r;
}

// On a non-increasing column the first match is the smallest k with d >= z[k]-eps (lower bound).
template <int M, bool FULL>
__device__ __forceinline__ LayerRes layer_search_path(const ZCol<M, FULL>& z, double d, int hint)
{
    const double eps = 1e-8;
    const int L = z.L;
    LayerRes r;
    if (hint >= 1 && hint <= L - 1) {
        r.layer = hint;
        z.pair(hint, r.top, r.bot);
        // hint is the first match iff it matches and hint-1 does not (or hint == 1)
        if (d >= r.bot - eps && d <= r.top + eps && (hint == 1 ? true : (d < r.top - eps))) return r;
    }
    r.top = 0.0; r.bot = 0.0;
    if (d > z(0) + eps) { r.layer = 0; return r; }
    int layer;
    if (d < z(L - 1) - eps) {
        layer = L - 1;
    } else {
        int lo = 1, hi = L - 1; // smallest k with d >= z[k]-eps; exists because d >= z[L-1]-eps
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (d >= z(mid) - eps) hi = mid;
            else lo = mid + 1;
        }
        layer = lo;
    }
    r.layer = layer;
    z.pair(layer, r.top, r.bot);
    return r;
}

// TBBKernel::CalcVelocity + CalcAttribute on the packed (vx,vy,vz,w) records, TK:128-164
template <int M, bool FULL = false>
__device__ __forceinline__ void gather_velw(const double4* __restrict__ velw, const voff_t (&vo)[M], const double (&w)[M], int nv,
                                            int layer, double& x, double& y, double& z, double& ww)
{
    x = 0.0; y = 0.0; z = 0.0; ww = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        if (FULL || i < nv) {
            const double4 v = ldg_d4(velw + vo[i] + layer);
            x += w[i] * v.x;
            y += w[i] * v.y;
            z += w[i] * v.z;
            ww += w[i] * v.w;
        }
    }
}

// levels `layer` (d*) and `layer - 1` (u*) of one snapshot in one pass over the vertices: one record pointer per
// vertex, the upper level at a fixed -32 B offset.  Every accumulator sums in vertex order exactly as
// gather_velw does, so the results are bit-equal to two gather_velw calls.
// wz: the snapshot's w slots hold zTop (SnapView::w_is_z): the vertical velocity is +0.0 there.
template <int M, bool FULL = false>
__device__ __forceinline__ void gather_velw_pair(const double4* __restrict__ velw, bool wz, const voff_t (&vo)[M], const double (&w)[M], int nv,
                                                 int layer, double& dx, double& dy, double& dz, double& dw,
                                                 double& ux, double& uy, double& uz, double& uw)
{
    dx = 0.0; dy = 0.0; dz = 0.0; dw = 0.0;
    ux = 0.0; uy = 0.0; uz = 0.0; uw = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        if (FULL || i < nv) {
            const double4* __restrict__ q = velw + (vo[i] + (voff_t)layer); // nV * L < 2^32 (checked at upload)
            const double4 d = ldg_d4(q);
            const double4 u = ldg_d4(q - 1);
            dx += w[i] * d.x;
            dy += w[i] * d.y;
            dz += w[i] * d.z;
            dw += w[i] * (wz ? 0.0 : d.w);
            ux += w[i] * u.x;
            uy += w[i] * u.y;
            uz += w[i] * u.z;
            uw += w[i] * (wz ? 0.0 : u.w);
        }
    }
}

template <int M, bool FULL = false>
__device__ __forceinline__ double gather_scalar(const double* __restrict__ a, const voff_t (&vo)[M], const double (&w)[M], int nv, int layer)
{
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i)
        if (FULL || i < nv) r += w[i] * a[vo[i] + layer];
    return r;
}

// MOPS_LENGTH(v) < 1e-12 (VK:845-852)
__device__ __forceinline__ bool tiny_len(double x, double y, double z) { return len3(x, y, z) < 1e-12; }

struct EvalOut {
    double hx, hy, hz; // horizontal velocity (XYZ)
    double vv;         // vertical velocity
    double a0, a1;     // pathline attributes
};

// calc_velocity_at (streamline), VK:740-872.  Returns ST_ALIVE or the reason of failure.
// FULL = true is the hexagon fast path (nv == M): it covers cells whose columns are monotone and points whose weight
// arithmetic stays inside the exact-sequence windows, and returns ST_GENERIC for everything else (the caller then runs the
// generic evaluation, FULL = false, which is the complete restatement).
template <int M, bool FULL = false>
__device__ __forceinline__ int eval_stream(const CellRec<M>* __restrict__ rec, const SnapView& s, bool mono, int L,
                                           const d3& p, double depth, int& hint, EvalOut& o)
{
    const int nv = FULL ? M : rec->nv;
    if (nv <= 0) return ST_BAD_SETUP;
    double w[M];
    bool wfinite = false;
    if (FULL) {
        if (!mono) return ST_GENERIC;
        if (!in_mesh<M, true>(rec, nv, p.x, p.y, p.z)) return ST_NOT_IN_CELL;
        hex_weights<M>(rec, p.x, p.y, p.z, w, wfinite);
        if (!wfinite) return ST_GENERIC;
    } else {
        if (!cell_weights<M, false>(rec, nv, p.x, p.y, p.z, w, wfinite)) return ST_NOT_IN_CELL;
    }
    voff_t vo[M];
#pragma unroll
    for (int i = 0; i < M; ++i) vo[i] = (FULL || i < nv) ? (voff_t)rec->vid[i] * (voff_t)L : 0u;

    LayerRes lr;
    if (FULL || (mono && wfinite)) lr = layer_search_stream<M, FULL>(ZCol<M, FULL>{s.ztop, w, vo, nv, L}, depth, hint);
    else lr = slow_layer_stream<M>(rec, s.ztop, L, p.x, p.y, p.z, depth);
    const int layer = lr.layer;
    const double ztop_up = lr.top, ztop_dn = lr.bot;
    hint = layer;
    // x = max(dn, min(depth, up)) with libstdc++ min/max semantics (VK:830-831)
    const double mn = (ztop_up < depth) ? ztop_up : depth;
    const double x = (ztop_dn < mn) ? mn : ztop_dn;
    const double denom = ztop_up - ztop_dn;
    if (fabs(denom) < 1e-12) return ST_BAD_COLUMN;
    const double t = (x - ztop_dn) / denom;

    double dx, dy, dz, dw, ux, uy, uz, uw;
    gather_velw_pair<M, FULL>(s.velw, s.w_is_z != 0, vo, w, nv, layer, dx, dy, dz, dw, ux, uy, uz, uw);
    if (tiny_len(dx, dy, dz) || tiny_len(ux, uy, uz)) return ST_ZERO_VELOCITY; // VK:845-847
    const double omt = 1.0 - t;
    o.hx = t * ux + omt * dx; // VK:849
    o.hy = t * uy + omt * dy;
    o.hz = t * uz + omt * dz;
    if (tiny_len(o.hx, o.hy, o.hz)) return ST_ZERO_VELOCITY;
    o.vv = t * uw + omt * dw; // VK:870 (levels `layer`, `layer-1` of vertVelocityTop)
    o.a0 = 0.0; o.a1 = 0.0;
    return ST_ALIVE;
}

// calc_velocity_at (pathline), VK:1124-1327: front and back interpolated separately (own layer
// search each, shared weights), blended with alpha; no zero-velocity reject.  sv[0] = front,
// sv[1] = back; the two snapshots go through ONE copy of the code (loops kept rolled) to
// keep the kernel's instruction footprint down.  FULL as in eval_stream.
template <int M, bool FULL = false>
__device__ __forceinline__ int eval_path(const CellRec<M>* __restrict__ rec, const SnapView* __restrict__ sv,
                                         bool mono_f, bool mono_b, int L, int attr_count, const d3& p, double depth,
                                         double alpha, int& hint_f, int& hint_b, EvalOut& o)
{
    const int nv = FULL ? M : rec->nv;
    if (nv <= 0) return ST_BAD_SETUP;
    double w[M];
    bool wfinite = false;
    if (FULL) {
        if (!(mono_f && mono_b)) return ST_GENERIC;
        if (!in_mesh<M, true>(rec, nv, p.x, p.y, p.z)) return ST_NOT_IN_CELL;
        hex_weights<M>(rec, p.x, p.y, p.z, w, wfinite);
        if (!wfinite) return ST_GENERIC;
    } else {
        if (!cell_weights<M, false>(rec, nv, p.x, p.y, p.z, w, wfinite)) return ST_NOT_IN_CELL;
    }
    voff_t vo[M];
#pragma unroll
    for (int i = 0; i < M; ++i) vo[i] = (FULL || i < nv) ? (voff_t)rec->vid[i] * (voff_t)L : 0u;

    // both layer searches first (VK:1182-1222) ...
    int lf = -1, lb = -1;
    double f_up = 0.0, f_dn = 0.0, b_up = 0.0, b_dn = 0.0;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const bool mono = s ? mono_b : mono_f;
        const int hint = s ? hint_b : hint_f;
        LayerRes r;
        if (FULL || (mono && wfinite)) r = layer_search_path<M, FULL>(ZCol<M, FULL>{sv[s].ztop, w, vo, nv, L}, depth, hint);
        else r = slow_layer_path<M>(rec, sv[s].ztop, L, p.x, p.y, p.z, depth);
        if (s == 0) { lf = r.layer; f_up = r.top; f_dn = r.bot; }
        else { lb = r.layer; b_up = r.top; b_dn = r.bot; }
    }
    if (lf == 0 || lb == 0) return ST_ABOVE_SURFACE; // the reference reads ztop[-1] here (N2)
    if (lf < 0 || lb < 0) return ST_BAD_COLUMN;
    hint_f = lf;
    hint_b = lb;

    // ... then denominators (front, back: VK:1229-1241) ...
    double mn = (f_up < depth) ? f_up : depth;
    const double x_front = (f_dn < mn) ? mn : f_dn;
    const double denom_front = f_up - f_dn;
    if (fabs(denom_front) < 1e-12) return ST_BAD_COLUMN;
    mn = (b_up < depth) ? b_up : depth;
    const double x_back = (b_dn < mn) ? mn : b_dn;
    const double denom_back = b_up - b_dn;
    if (fabs(denom_back) < 1e-12) return ST_BAD_COLUMN;
    // the two quotients are independent: one group (A/B on B200: 1.3 % of the kernel)
    double tq[2] = {x_front - f_dn, x_back - b_dn};
    const double td[2] = {denom_front, denom_back};
    div_group<2>(tq, td);
    const double t_front = tq[0], t_back = tq[1];

    // ... then the gathers and the alpha blend (VK:1243-1324)
    const double oma = 1.0 - alpha;
    double ffx = 0.0, ffy = 0.0, ffz = 0.0, ffw = 0.0, fa0 = 0.0, fa1 = 0.0;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const int layer = s ? lb : lf;
        const double t = s ? t_back : t_front;
        const double omt = 1.0 - t;
        double dx, dy, dz, dw, ux, uy, uz, uw;
        gather_velw_pair<M, FULL>(sv[s].velw, sv[s].w_is_z != 0, vo, w, nv, layer, dx, dy, dz, dw, ux, uy, uz, uw);
        const double vx = t * ux + omt * dx, vy = t * uy + omt * dy, vz = t * uz + omt * dz;
        const double vw = t * uw + omt * dw;
        double a0 = 0.0, a1 = 0.0;
        if (attr_count >= 1) {
            const double ad = gather_scalar<M, FULL>(sv[s].attr0, vo, w, nv, layer), au = gather_scalar<M, FULL>(sv[s].attr0, vo, w, nv, layer - 1);
            a0 = t * au + omt * ad;
        }
        if (attr_count >= 2) {
            const double ad = gather_scalar<M, FULL>(sv[s].attr1, vo, w, nv, layer), au = gather_scalar<M, FULL>(sv[s].attr1, vo, w, nv, layer - 1);
            a1 = t * au + omt * ad;
        }
        if (s == 0) {
            ffx = vx; ffy = vy; ffz = vz; ffw = vw; fa0 = a0; fa1 = a1;
        } else {
            o.hx = alpha * vx + oma * ffx; // VK:1259
            o.hy = alpha * vy + oma * ffy;
            o.hz = alpha * vz + oma * ffz;
            o.vv = alpha * vw + oma * ffw; // VK:1286
            o.a0 = (attr_count >= 1) ? alpha * a0 + oma * fa0 : 0.0;
            o.a1 = (attr_count >= 2) ? alpha * a1 + oma * fa1 : 0.0;
        }
    }
    return ST_ALIVE;
}

// ---- greedy nearest-centre walk (thread-serial form) -------------------------------------
// On a Voronoi/Delaunay mesh every cell that is not the nearest generator to q has a
// cellsOnCell neighbour strictly closer to q, so steepest descent on dist2 ends at the
// exact nearest centre = what the reference's nanoflann 1-NN returns
// (src/Core/MPASOGrid.cpp:287-313).
template <int M>
__device__ __forceinline__ int walk_nearest(const CellRec<M>* __restrict__ rec, const double4* __restrict__ c4, int start,
                                            double qx, double qy, double qz)
{
    int cur = start;
    double4 c = c4[cur];
    double dcur = dist2(qx, qy, qz, c.x, c.y, c.z);
    for (int it = 0; it < (1 << 22); ++it) {
        const CellRec<M>* r = rec + cur;
        const int nv = r->nv;
        int best = cur;
        double dbest = dcur;
#pragma unroll
        for (int k = 0; k < M; ++k) {
            if (k < nv) {
                const int nb = r->nbr[k];
                if (nb >= 0) {
                    const double4 cc = c4[nb];
                    const double d = dist2(qx, qy, qz, cc.x, cc.y, cc.z);
                    if (d < dbest) {
                        dbest = d;
                        best = nb;
                    }
                }
            }
        }
        if (best == cur) break;
        cur = best;
        dcur = dbest;
    }
    return cur;
}

// ---- exact 1-NN over the cell centres for meshes with removed (land) cells ----------------------
// On a culled mesh the greedy cellsOnCell walk can stop at a local minimum behind a coast, and a query on
// land has no containing cell at all.  A host-built kd-tree (median splits, implicit layout: the node of
// range [lo,hi) is element (lo+hi)/2, children [lo,mid) and (mid,hi)) gives the same answer as the
// reference's nanoflann search (src/Core/MPASOGrid.cpp:287-313) for those queries.
struct KdView {
    const double4* pts;       // tree order; .w carries the (internal) cell id bits
    const unsigned char* dim; // split axis of each node
    int n;                    // 0: closed mesh, the walk alone is exact
};

__device__ __noinline__ int kd_nearest(const KdView kd, double qx, double qy, double qz, int best, double dbest)
{
    int s_lo[40], s_hi[40];
    double s_b[40];
    int sp = 0;
    s_lo[0] = 0; s_hi[0] = kd.n; s_b[0] = 0.0; sp = 1;
    while (sp > 0) {
        --sp;
        int lo = s_lo[sp], hi = s_hi[sp];
        if (!(s_b[sp] < dbest)) continue; // nothing beyond that plane can be strictly closer
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double4 p = kd.pts[mid];
            const double d = dist2(qx, qy, qz, p.x, p.y, p.z);
            const int id = (int)__double_as_longlong(p.w);
            if (d < dbest || (d == dbest && id < best)) { dbest = d; best = id; }
            const int ax = kd.dim[mid];
            const double diff = (ax == 0) ? qx - p.x : (ax == 1) ? qy - p.y : qz - p.z;
            int flo, fhi;
            if (diff < 0.0) { flo = mid + 1; fhi = hi; hi = mid; }
            else            { flo = lo; fhi = mid; lo = mid + 1; }
            const double b = diff * diff;
            if (flo < fhi && b < dbest && sp < 40) { s_lo[sp] = flo; s_hi[sp] = fhi; s_b[sp] = b; ++sp; }
        }
    }
    return best;
}

// direction -> cube-map bucket (start guess of the walk; float precision is enough)
__device__ __forceinline__ int cube_bucket(double x, double y, double z, int F)
{
    const double axx = fabs(x), ayy = fabs(y), azz = fabs(z);
    int face;
    double ma, u, v;
    if (axx >= ayy && axx >= azz) { face = x >= 0 ? 0 : 1; ma = axx; u = y; v = z; }
    else if (ayy >= azz)          { face = y >= 0 ? 2 : 3; ma = ayy; u = x; v = z; }
    else                          { face = z >= 0 ? 4 : 5; ma = azz; u = x; v = y; }
    const float inv = 1.0f / (float)ma;
    int i = (int)(((float)u * inv * 0.5f + 0.5f) * (float)F);
    int j = (int)(((float)v * inv * 0.5f + 0.5f) * (float)F);
    i = min(max(i, 0), F - 1);
    j = min(max(j, 0), F - 1);
    return (face * F + i) * F + j;
}

// centre direction of bucket (face, i, j)
__device__ __forceinline__ void cube_center(int face, int i, int j, int F, double& x, double& y, double& z)
{
    const double u = ((double)i + 0.5) / (double)F * 2.0 - 1.0;
    const double v = ((double)j + 0.5) / (double)F * 2.0 - 1.0;
    const double s = (face & 1) ? -1.0 : 1.0;
    if (face < 2)      { x = s; y = u; z = v; }
    else if (face < 4) { x = u; y = s; z = v; }
    else               { x = u; y = v; z = s; }
}

// point location as the reference defines it (nearest cell centre): cube-map start + greedy walk; on a
// culled mesh the answer is accepted only when the point lies inside that cell's polygon, otherwise the
// kd-tree decides.
template <int M>
__device__ __forceinline__ int locate_cell(const CellRec<M>* __restrict__ rec, const double4* __restrict__ c4, const int* __restrict__ cube,
                                           int F, const KdView kd, double qx, double qy, double qz)
{
    int c = walk_nearest<M>(rec, c4, cube[cube_bucket(qx, qy, qz, F)], qx, qy, qz);
    if (kd.n > 0 && !in_mesh<M>(rec + c, rec[c].nv, qx, qy, qz)) {
        const double4 cc = c4[c];
        c = kd_nearest(kd, qx, qy, qz, c, dist2(qx, qy, qz, cc.x, cc.y, cc.z));
    }
    return c;
}

} // namespace mops
