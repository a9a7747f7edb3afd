/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product.
 *
 * mops_oracle.c : a plain scalar C restatement ("Tier B") of the reference's hot path
 * (YosefQiu/MOPS, TBB/CPU backend), one function per reference function, each citing the
 * file:line it follows.  Paths are relative to the reference root; "VK" abbreviates
 * src/CPU/TBB/Kernel/MPASOVisualizerKernels.cpp and "TK" src/CPU/TBB/Kernel/TBBKernel.h.
 *
 * PARITY PINNING.  The reference's own tests hold no golden vector for this path
 * (SURVEY.md section 4): the only pins are test/test_gaussian.cpp (3x3 solve) and
 * test/test_trajector.cpp (NaN trimming).  This restatement is therefore pinned against
 * OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref/libmops_ref.so is the reference's TBB
 * backend compiled unmodified (oracle/build_ref.sh); tests/test_oracle_vs_ref.py requires
 * every position / velocity / pixel / prepared array of this file to be BIT-IDENTICAL to
 * it, and tests/golden/ holds committed vectors generated from it
 * (tests/golden/make_golden.py).  What this file adds over the compiled reference is the
 * per-step cell-id log and the raw (pre-assembly) buffers, which the reference does not
 * expose, and that it travels to the GPU box as source.
 *
 * Arithmetic contract: compiled with `gcc -O2 -ffp-contract=off`, no -march, no
 * -ffast-math -- the same evaluation as the reference's `g++ -O2` x86-64 build (baseline
 * x86-64 has no FMA, so every product is rounded before the add).  Expressions are kept
 * textually in the reference's association order; do not "simplify" them.
 *
 * Deliberately NOT restated: the 20-entry over-read of a
 * maxEdges-wide row (TK:64-66), ztop[-1] read of the pathline above-surface branch
 * (VK:1188-1190,1225-1227: reported as status ORC_ST_ABOVE_SURFACE instead), the
 * by-value SetPixel, the disk cache, 32-bit slot indexing.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_MAX_VERTEX_NUM 20          /* VK:741, VK:1125 */
#define ORC_MAX_VERTICAL_LEVEL_NUM 100 /* VK:742, VK:1126 */
#define ORC_MAX_CELL_NEIGHBOR_NUM 21   /* VK:875 */

/* per-particle termination status (reported; the reference just `return`s) */
enum {
    ORC_ST_ALIVE = 0,
    ORC_ST_BAD_CELL = 1,      /* cell id out of range                       VK:895,903    */
    ORC_ST_NOT_IN_CELL = 2,   /* IsInMesh false at some RK stage            VK:753-756    */
    ORC_ST_BAD_COLUMN = 3,    /* |denom| < 1e-12 or invalid vertex          VK:833,777    */
    ORC_ST_ZERO_VELOCITY = 4, /* |v| < 1e-12 reject (streamline only)       VK:845-852    */
    ORC_ST_ABOVE_SURFACE = 5, /* pathline local_layer == 0 (N2, not restated) VK:1188-1190 */
    ORC_ST_BAD_SETUP = 6
};

typedef struct { double x, y, z; } v3;

static inline v3 v3_make(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline v3 v3_ld(const double* p, int64_t i) { return v3_make(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
static inline void v3_st(double* p, int64_t i, v3 a) { p[3 * i] = a.x; p[3 * i + 1] = a.y; p[3 * i + 2] = a.z; }
/* cy::Vec3d operators, src/Utils/CPUCommon/cyVector.h:361-393 */
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, double s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_div(v3 a, double s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline v3 v3_cross(v3 a, v3 p) { return v3_make(a.y * p.z - a.z * p.y, a.z * p.x - a.x * p.z, a.x * p.y - a.y * p.x); }
static inline double v3_dot(v3 a, v3 p) { return a.x * p.x + a.y * p.y + a.z * p.z; }
/* MOPS_LENGTH, src/Utils/BackendCompat.hpp (USE_TBB branch) */
static inline double v3_len(v3 v) { return sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }

typedef struct {
    int n_cells, n_vertices, max_edges, n_levels;
    const double* cell_xyz;          /* [n_cells][3]   */
    const double* vertex_xyz;        /* [n_vertices][3] */
    const int32_t* vertices_on_cell; /* [n_cells][max_edges] 1-based, 0 pad */
    const int32_t* cells_on_cell;    /* [n_cells][max_edges] 1-based, 0 pad */
    const int32_t* n_edges_on_cell;  /* [n_cells] */
} orc_mesh;

typedef struct {
    const double* ztop_v;  /* [n_vertices][L]     cellVertexZTop_vec          */
    const double* vel_v;   /* [n_vertices][L][3]  cellVertexVelocity_vec      */
    const double* w_v;     /* [n_vertices][L+1]   cellVertexVertVelocity_vec  */
    const double* attr0_v; /* [n_vertices][L] or NULL  (mDoubleAttributes_CtoV, std::map order) */
    const double* attr1_v;
} orc_fields;

/* ------------------------------------------------------------------------------------- */
/* Geometry helpers                                                                       */
/* ------------------------------------------------------------------------------------- */

/* Interpolator::triangle_area, src/Utils/Interpolation.hpp:95-110 */
static double triangle_area(v3 a, v3 b, v3 c)
{
    v3 edge1 = v3_make(b.x - a.x, b.y - a.y, b.z - a.z);
    v3 edge2 = v3_make(c.x - a.x, c.y - a.y, c.z - a.z);
    v3 cp = v3_make(edge1.y * edge2.z - edge1.z * edge2.y,
                    edge1.z * edge2.x - edge1.x * edge2.z,
                    edge1.x * edge2.y - edge1.y * edge2.x);
    return sqrt(cp.x * cp.x + cp.y * cp.y + cp.z * cp.z) / 2.0f;
}

/* Interpolator::CalcPolygonWachspress (pointer overload), src/Utils/Interpolation.hpp:137-165 */
static void wachspress(v3 p, const v3* poly, double* weights, int N)
{
    for (int i = 0; i < N; i++) weights[i] = 0.0;
    double sumweights = 0.0;
    double A_i, A_iplus1, B;
    A_iplus1 = triangle_area(poly[N - 1], poly[0], p);
    for (int i = 0; i < N; i++) {
        A_i = A_iplus1;
        A_iplus1 = triangle_area(poly[i], poly[(i + 1) % N], p);
        B = triangle_area(poly[(i - 1 + N) % N], poly[i], poly[(i + 1) % N]);
        weights[i] = B / (A_i * A_iplus1);
        sumweights += weights[i];
    }
    double recp = 1.0 / sumweights;
    for (int i = 0; i < N; i++) weights[i] *= recp;
}

/* Interpolator::calcTriangleBarycentric, src/Utils/Interpolation.hpp:79-93 */
static void triangle_barycentric(v3 p, v3 t0, v3 t1, v3 t2, double* u, double* v, double* w)
{
    v3 v0 = v3_sub(t1, t0);
    v3 v1 = v3_sub(t2, t0);
    v3 v2 = v3_sub(p, t0);
    double d00 = v3_dot(v0, v0);
    double d01 = v3_dot(v0, v1);
    double d11 = v3_dot(v1, v1);
    double d20 = v3_dot(v2, v0);
    double d21 = v3_dot(v2, v1);
    double denom = d00 * d11 - d01 * d01;
    *v = (d11 * d20 - d01 * d21) / denom;
    *w = (d00 * d21 - d01 * d20) / denom;
    *u = 1.0 - *v - *w;
}

/* GeoConverter::convertENUVelocityToXYZ, src/Utils/GeoConverter.hpp:225-250 */
static v3 enu_to_xyz(v3 p, double Uzon, double Umer, double Uup)
{
    v3 out;
    if (p.x == 0.0 && p.y == 0.0) {
        out.x = 0.0; out.y = 0.0; out.z = Uup;
        return out;
    }
    double Rxy = sqrt(p.x * p.x + p.y * p.y);
    double Rxyz = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    double slon = p.y / Rxy;
    double clon = p.x / Rxy;
    double slat = p.z / Rxyz;
    double clat = Rxy / Rxyz;
    out.x = -slon * Uzon - slat * clon * Umer + clon * clat * Uup;
    out.y = clon * Uzon - slat * slon * Umer + slon * clat * Uup;
    out.z = clat * Umer + slat * Uup;
    return out;
}

/* GeoConverter::convertXYZVelocityToENU, src/Utils/GeoConverter.hpp:200-223 */
static void xyz_to_enu(v3 p, v3 vel, double* Uzon, double* Umer)
{
    if (p.x == 0.0 && p.y == 0.0) { *Uzon = 0.0; *Umer = 0.0; return; }
    double Rxy = sqrt(p.x * p.x + p.y * p.y);
    double Rxyz = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
    double slon = p.y / Rxy;
    double clon = p.x / Rxy;
    double slat = p.z / Rxyz;
    double clat = Rxy / Rxyz;
    *Uzon = -slon * vel.x + clon * vel.y;
    *Umer = -slat * (clon * vel.x + slon * vel.y) + clat * vel.z;
}

/* ------------------------------------------------------------------------------------- */
/* a17: preprocessing that produces the gathered arrays                                   */
/* ------------------------------------------------------------------------------------- */

/* MPASOSolution::calcCellCenterZtop, bottomDepth branch, src/Core/MPASOSolution.cpp:565-577 */
void orc_cell_ztop(int n_cells, int L, const double* layer_thickness, const double* bottom_depth, double* ztop_c)
{
    for (int64_t i = 0; i < n_cells; ++i) {
        double z = -bottom_depth[i];
        for (int k = L - 1; k >= 0; --k) {
            z += layer_thickness[i * L + k];
            ztop_c[i * L + k] = z;
        }
    }
    for (int64_t i = 0; i < (int64_t)n_cells * L; ++i) ztop_c[i] *= 1.0; /* :598-601 */
}

/* CalcCellCenterVelocityByZM, src/CPU/TBB/MPASOSolutionTBB.cpp:108-129 */
void orc_cell_velocity_zm(int n_cells, int L, const double* cell_xyz, const double* zonal, const double* merid, double* vel_c)
{
    for (int64_t idx = 0; idx < (int64_t)n_cells * L; ++idx) {
        int64_t c = idx / L;
        v3 v = enu_to_xyz(v3_ld(cell_xyz, c), zonal[idx], merid[idx], 0.0);
        v3_st(vel_c, idx, v);
    }
}

/* Shared body of CalcCellVertexZtop (:9-55), CalcCellCenterToVertex (:57-106, clamp0),
 * CalcCellVertexVertVelocity (:320-366, levels = L+1), src/CPU/TBB/MPASOSolutionTBB.cpp.
 * Boundary test kept as written: (id-1 as size_t) > n_cells + 1  -> whole vertex = 0. */
void orc_cell_to_vertex_scalar(int n_cells, int n_vertices, int levels, const double* cell_xyz, const double* vertex_xyz,
                               const int32_t* cells_on_vertex, const double* cell_val, int clamp0, double* vert_val)
{
    for (int64_t v = 0; v < n_vertices; ++v) {
        size_t id[3];
        int boundary = 0;
        for (int t = 0; t < 3; ++t) {
            id[t] = (size_t)cells_on_vertex[3 * v + t] - 1;
            if (id[t] > (size_t)(n_cells + 1)) boundary = 1;
        }
        double bu = 0, bv = 0, bw = 0;
        if (!boundary) {
            triangle_barycentric(v3_ld(vertex_xyz, v), v3_ld(cell_xyz, (int64_t)id[0]), v3_ld(cell_xyz, (int64_t)id[1]),
                                 v3_ld(cell_xyz, (int64_t)id[2]), &bu, &bv, &bw);
        }
        for (int k = 0; k < levels; ++k) {
            double out = 0.0;
            if (!boundary) {
                double a = cell_val[(int64_t)levels * (int64_t)id[0] + k];
                double b = cell_val[(int64_t)levels * (int64_t)id[1] + k];
                double c = cell_val[(int64_t)levels * (int64_t)id[2] + k];
                out = bu * a + bv * b + bw * c;
                if (clamp0 && out < 0.0) out = 0.0;
            }
            vert_val[v * levels + k] = out;
        }
    }
}

/* CalcCellVertexVelocity, src/CPU/TBB/MPASOSolutionTBB.cpp:270-318 */
void orc_cell_to_vertex_vec3(int n_cells, int n_vertices, int levels, const double* cell_xyz, const double* vertex_xyz,
                             const int32_t* cells_on_vertex, const double* cell_val, double* vert_val)
{
    for (int64_t v = 0; v < n_vertices; ++v) {
        size_t id[3];
        int boundary = 0;
        for (int t = 0; t < 3; ++t) {
            id[t] = (size_t)cells_on_vertex[3 * v + t] - 1;
            if (id[t] > (size_t)(n_cells + 1)) boundary = 1;
        }
        double bu = 0, bv = 0, bw = 0;
        if (!boundary) {
            triangle_barycentric(v3_ld(vertex_xyz, v), v3_ld(cell_xyz, (int64_t)id[0]), v3_ld(cell_xyz, (int64_t)id[1]),
                                 v3_ld(cell_xyz, (int64_t)id[2]), &bu, &bv, &bw);
        }
        for (int k = 0; k < levels; ++k) {
            v3 out = v3_make(0.0, 0.0, 0.0);
            if (!boundary) {
                v3 a = v3_ld(cell_val, (int64_t)levels * (int64_t)id[0] + k);
                v3 b = v3_ld(cell_val, (int64_t)levels * (int64_t)id[1] + k);
                v3 c = v3_ld(cell_val, (int64_t)levels * (int64_t)id[2] + k);
                out = v3_add(v3_add(v3_mul(a, bu), v3_mul(b, bv)), v3_mul(c, bw));
            }
            v3_st(vert_val, v * levels + k, out);
        }
    }
}

/* The chain MOPSApp::addSol runs for one snapshot, src/Core/MOPSApp.cpp:100-130.
 * Scratch (ztop_c [nC*L], vel_c [nC*L*3]) may be NULL (allocated here). */
int orc_prepare_snapshot(int n_cells, int n_vertices, int L, const double* cell_xyz, const double* vertex_xyz,
                         const int32_t* cells_on_vertex, const double* zonal, const double* merid,
                         const double* layer_thickness, const double* bottom_depth, const double* vert_vel_top,
                         double* ztop_v, double* vel_v, double* w_v, double* ztop_c_out, double* vel_c_out)
{
    double* ztop_c = ztop_c_out ? ztop_c_out : (double*)malloc(sizeof(double) * (size_t)n_cells * L);
    double* vel_c = vel_c_out ? vel_c_out : (double*)malloc(sizeof(double) * (size_t)n_cells * L * 3);
    if (!ztop_c || !vel_c) return -1;
    orc_cell_ztop(n_cells, L, layer_thickness, bottom_depth, ztop_c);
    orc_cell_to_vertex_scalar(n_cells, n_vertices, L, cell_xyz, vertex_xyz, cells_on_vertex, ztop_c, 0, ztop_v);
    orc_cell_velocity_zm(n_cells, L, cell_xyz, zonal, merid, vel_c);
    orc_cell_to_vertex_vec3(n_cells, n_vertices, L, cell_xyz, vertex_xyz, cells_on_vertex, vel_c, vel_v);
    orc_cell_to_vertex_scalar(n_cells, n_vertices, L + 1, cell_xyz, vertex_xyz, cells_on_vertex, vert_vel_top, 0, w_v);
    if (!ztop_c_out) free(ztop_c);
    if (!vel_c_out) free(vel_c);
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* a1/a2: exact nearest cell centre (what nanoflann returns, src/Core/MPASOGrid.cpp:287-313) */
/* ------------------------------------------------------------------------------------- */

/* squared distance exactly as nanoflann's L2_Adaptor tail loop accumulates it for dim = 3
 * (src/Utils/nanoflann.hpp, metric_L2): result += (a[i]-b[i])^2, i = 0,1,2 */
static inline double dist2(const double* q, const double* c)
{
    double r = 0.0;
    double d0 = q[0] - c[0]; r += d0 * d0;
    double d1 = q[1] - c[1]; r += d1 * d1;
    double d2 = q[2] - c[2]; r += d2 * d2;
    return r;
}

typedef struct {
    int nb;
    double lo[3], inv_h, h;
    int32_t* start; /* [nb^3+1] */
    int32_t* items; /* [n_cells] */
} orc_bins;

static int bin_of(const orc_bins* b, double v, int axis)
{
    int i = (int)floor((v - b->lo[axis]) * b->inv_h);
    if (i < 0) i = 0;
    if (i >= b->nb) i = b->nb - 1;
    return i;
}

static orc_bins* bins_build(int n_cells, const double* cell_xyz)
{
    orc_bins* b = (orc_bins*)calloc(1, sizeof(orc_bins));
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int64_t i = 0; i < n_cells; ++i)
        for (int a = 0; a < 3; ++a) {
            double v = cell_xyz[3 * i + a];
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    double ext = 0;
    for (int a = 0; a < 3; ++a) { b->lo[a] = lo[a]; if (hi[a] - lo[a] > ext) ext = hi[a] - lo[a]; }
    /* points live on a sphere: ~n_cells/(pi*nb^2) per occupied bin; aim for a handful */
    int nb = (int)floor(sqrt((double)n_cells / 6.0));
    if (nb < 1) nb = 1;
    if (nb > 256) nb = 256;
    b->nb = nb;
    b->h = ext / nb * (1.0 + 1e-9) + 1e-30;
    b->inv_h = 1.0 / b->h;
    int64_t nbin = (int64_t)nb * nb * nb;
    b->start = (int32_t*)calloc((size_t)nbin + 1, sizeof(int32_t));
    b->items = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_cells);
    for (int64_t i = 0; i < n_cells; ++i) {
        int64_t id = ((int64_t)bin_of(b, cell_xyz[3 * i], 0) * nb + bin_of(b, cell_xyz[3 * i + 1], 1)) * nb + bin_of(b, cell_xyz[3 * i + 2], 2);
        b->start[id + 1]++;
    }
    for (int64_t i = 0; i < nbin; ++i) b->start[i + 1] += b->start[i];
    int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)nbin);
    memcpy(fill, b->start, sizeof(int32_t) * (size_t)nbin);
    for (int64_t i = 0; i < n_cells; ++i) {
        int64_t id = ((int64_t)bin_of(b, cell_xyz[3 * i], 0) * nb + bin_of(b, cell_xyz[3 * i + 1], 1)) * nb + bin_of(b, cell_xyz[3 * i + 2], 2);
        b->items[fill[id]++] = (int32_t)i;
    }
    free(fill);
    return b;
}

static void bins_free(orc_bins* b) { if (b) { free(b->start); free(b->items); free(b); } }

/* exact 1-NN: grow a cube of bins until the best distance is provably inside it;
 * ties resolve to the lowest cell index (a tie needs a query exactly on a bisector). */
static int32_t bins_nearest(const orc_bins* b, int n_cells, const double* cell_xyz, const double* q)
{
    if (!(isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2]))) return -1;
    int c[3];
    for (int a = 0; a < 3; ++a) c[a] = bin_of(b, q[a], a);
    int nb = b->nb;
    for (int ring = 1; ring <= nb; ++ring) {
        double best = DBL_MAX;
        int32_t best_i = -1;
        int lo[3], hi[3];
        for (int a = 0; a < 3; ++a) { lo[a] = c[a] - ring; if (lo[a] < 0) lo[a] = 0; hi[a] = c[a] + ring; if (hi[a] >= nb) hi[a] = nb - 1; }
        for (int i = lo[0]; i <= hi[0]; ++i)
            for (int j = lo[1]; j <= hi[1]; ++j)
                for (int k = lo[2]; k <= hi[2]; ++k) {
                    int64_t id = ((int64_t)i * nb + j) * nb + k;
                    for (int32_t s = b->start[id]; s < b->start[id + 1]; ++s) {
                        int32_t ci = b->items[s];
                        double d = dist2(q, cell_xyz + 3 * (int64_t)ci);
                        if (d < best || (d == best && ci < best_i)) { best = d; best_i = ci; }
                    }
                }
        /* everything outside the scanned cube is at least this far from q */
        double guard = DBL_MAX;
        for (int a = 0; a < 3; ++a) {
            double dl = q[a] - (b->lo[a] + (c[a] - ring) * b->h);
            double dh = (b->lo[a] + (c[a] + ring + 1) * b->h) - q[a];
            if (c[a] - ring > 0 && dl < guard) guard = dl;
            if (c[a] + ring < nb - 1 && dh < guard) guard = dh;
        }
        if (best_i >= 0 && (guard == DBL_MAX || best < guard * guard * (1.0 - 1e-9))) return best_i;
    }
    /* brute force (only reachable for far-away queries) */
    double best = DBL_MAX;
    int32_t best_i = -1;
    for (int32_t ci = 0; ci < n_cells; ++ci) {
        double d = dist2(q, cell_xyz + 3 * (int64_t)ci);
        if (d < best) { best = d; best_i = ci; }
    }
    return best_i;
}

int orc_locate(int64_t n, const double* xyz, int n_cells, const double* cell_xyz, int32_t* cell_out)
{
    orc_bins* b = bins_build(n_cells, cell_xyz);
    for (int64_t i = 0; i < n; ++i) cell_out[i] = bins_nearest(b, n_cells, cell_xyz, xyz + 3 * i);
    bins_free(b);
    return 0;
}

int orc_locate_bruteforce(int64_t n, const double* xyz, int n_cells, const double* cell_xyz, int32_t* cell_out)
{
    for (int64_t i = 0; i < n; ++i) {
        double best = DBL_MAX;
        int32_t best_i = -1;
        for (int32_t ci = 0; ci < n_cells; ++ci) {
            double d = dist2(xyz + 3 * i, cell_xyz + 3 * (int64_t)ci);
            if (d < best) { best = d; best_i = ci; }
        }
        cell_out[i] = best_i;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* a3/a4/a8/a10/a11: TBBKernel helpers                                                    */
/* ------------------------------------------------------------------------------------- */

/* TBBKernel::IsInMesh, TK:21-54 */
static int is_in_mesh(const orc_mesh* m, int cell_id, v3 p)
{
    if (!isfinite(p.x) || !isfinite(p.y) || !isfinite(p.z)) return 0;
    int nv = m->n_edges_on_cell[cell_id];
    if (nv == 0) return 0;
    for (int k = 0; k < nv; ++k) {
        int64_t a_idx = (int64_t)m->vertices_on_cell[(int64_t)cell_id * m->max_edges + k] - 1;
        int64_t b_idx = (int64_t)m->vertices_on_cell[(int64_t)cell_id * m->max_edges + ((k + 1) % nv)] - 1;
        v3 a = v3_ld(m->vertex_xyz, a_idx);
        v3 b = v3_ld(m->vertex_xyz, b_idx);
        v3 surface_normal = v3_cross(a, b);
        double direction = v3_dot(surface_normal, p);
        if (direction < 0.0) return 0;
    }
    return 1;
}

/* TBBKernel::GetCellNeighborsIdx, TK:74-101 (neighbours in cellsOnCell order, self last) */
static void get_cell_neighbors(const orc_mesh* m, int cell_id, int nv, int* neig)
{
    const int VLA = ORC_MAX_CELL_NEIGHBOR_NUM;
    if (nv > VLA) return;
    neig[0] = cell_id;
    int copyN = nv;
    if (copyN > VLA - 1) copyN = VLA - 1;
    for (int k = 0; k < copyN; ++k) {
        int nid1 = (int)m->cells_on_cell[(int64_t)cell_id * m->max_edges + k];
        neig[k] = nid1 - 1;
    }
    neig[copyN] = cell_id;
    for (int k = copyN + 1; k < VLA; ++k) neig[k] = -1;
}

/* cell relocation, VK:903-921 (strict <, neighbour order, one ring) */
static int relocate(const orc_mesh* m, int cell_id, int* neig, v3 pos)
{
    int nv = m->n_edges_on_cell[cell_id];
    double min_len = DBL_MAX;
    for (int n = 0; n < nv + 1; ++n) {
        int cid = neig[n];
        if (cid < 0 || cid >= m->n_cells) continue;
        double len = v3_len(v3_sub(v3_ld(m->cell_xyz, cid), pos));
        if (len < min_len) { min_len = len; cell_id = cid; }
    }
    nv = m->n_edges_on_cell[cell_id];
    get_cell_neighbors(m, cell_id, nv, neig);
    return cell_id;
}

/* TBBKernel::CalcVelocity, TK:128-145 */
static v3 calc_velocity(const int64_t* vidx, const double* w, int nv, int L, int layer, const double* vel_v)
{
    v3 r = v3_make(0.0, 0.0, 0.0);
    for (int i = 0; i < nv; ++i) {
        v3 vel = v3_ld(vel_v, vidx[i] * L + layer);
        r.x += w[i] * vel.x;
        r.y += w[i] * vel.y;
        r.z += w[i] * vel.z;
    }
    return r;
}

/* TBBKernel::CalcAttribute, TK:147-164 */
static double calc_attribute(const int64_t* vidx, const double* w, int nv, int levels, int layer, const double* attr)
{
    double r = 0.0;
    for (int i = 0; i < nv; ++i) r += w[i] * attr[vidx[i] * levels + layer];
    return r;
}

/* TBBKernel::CalcRotationAxis, TK:166-173 */
static v3 rotation_axis(v3 position, v3 velocity)
{
    v3 axis;
    axis.x = position.y * velocity.z - position.z * velocity.y;
    axis.y = position.z * velocity.x - position.x * velocity.z;
    axis.z = position.x * velocity.y - position.y * velocity.x;
    return axis;
}

/* TBBKernel::CalcPositionAfterRotation, TK:175-204 */
static v3 position_after_rotation(v3 position, v3 axis, double theta_rad)
{
    const double cosTheta = cos(theta_rad);
    const double sinTheta = sin(theta_rad);
    const double axis_len = v3_len(axis);
    if (axis_len <= 1e-12) return position;
    v3 u;
    u.x = axis.x / axis_len;
    u.y = axis.y / axis_len;
    u.z = axis.z / axis_len;
    v3 rotated;
    rotated.x = (cosTheta + u.x * u.x * (1.0 - cosTheta)) * position.x +
        (u.x * u.y * (1.0 - cosTheta) - u.z * sinTheta) * position.y +
        (u.x * u.z * (1.0 - cosTheta) + u.y * sinTheta) * position.z;
    rotated.y = (u.y * u.x * (1.0 - cosTheta) + u.z * sinTheta) * position.x +
        (cosTheta + u.y * u.y * (1.0 - cosTheta)) * position.y +
        (u.y * u.z * (1.0 - cosTheta) - u.x * sinTheta) * position.z;
    rotated.z = (u.z * u.x * (1.0 - cosTheta) - u.y * sinTheta) * position.x +
        (u.z * u.y * (1.0 - cosTheta) + u.x * sinTheta) * position.y +
        (cosTheta + u.z * u.z * (1.0 - cosTheta)) * position.z;
    return rotated;
}

/* advect_on_sphere, VK:729-738 */
static v3 advect_on_sphere(v3 pos, v3 vel, double dt_local)
{
    const double rr = v3_len(pos);
    const double speed_local = v3_len(vel);
    if (rr < 1e-12 || speed_local < 1e-12) return pos;
    v3 axis = rotation_axis(pos, vel);
    const double theta = (speed_local * dt_local) / rr;
    return position_after_rotation(pos, axis, theta);
}

/* Gather of vertex ids / positions + Wachspress weights shared by every evaluation
 * (TK:56-72 without the over-read, TK:103-126, Interpolation.hpp:137-165). */
static int cell_weights(const orc_mesh* m, int cell_id, v3 pos, int nv, int64_t* vidx, double* w)
{
    v3 vpos[ORC_MAX_VERTEX_NUM];
    for (int k = 0; k < nv; ++k) {
        vidx[k] = (int64_t)m->vertices_on_cell[(int64_t)cell_id * m->max_edges + k] - 1;
        vpos[k] = v3_ld(m->vertex_xyz, vidx[k]);
    }
    wachspress(pos, vpos, w, nv);
    return 0;
}

/* zTop column at a point + monotone fix-up, VK:772-789 */
static int point_column(const orc_mesh* m, const int64_t* vidx, const double* w, int nv, const double* ztop_v, double* col)
{
    const int L = m->n_levels;
    for (int k = 0; k < L; ++k) {
        double z = 0.0;
        for (int i = 0; i < nv; ++i) {
            const int vid = (int)vidx[i];
            if (vid < 0 || vid >= m->n_vertices) return -1;
            z += w[i] * ztop_v[(int64_t)vid * L + k];
        }
        col[k] = z;
    }
    for (int k = 1; k < L; ++k)
        if (col[k] > col[k - 1]) col[k] = col[k - 1] - 1e-9;
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* a9: calc_velocity_at (streamline), VK:740-872                                          */
/* ------------------------------------------------------------------------------------- */
typedef struct { v3 h_vel; double v_vel; v3 attr; int ok; int why; int layer; } vel_state;

static vel_state fail_state(int why)
{
    vel_state s;
    s.h_vel = v3_make(0.0, 0.0, 0.0); s.v_vel = 0.0; s.attr = v3_make(0.0, 0.0, 0.0); s.ok = 0; s.why = why; s.layer = -1;
    return s;
}

static vel_state calc_velocity_at_stream(const orc_mesh* m, const orc_fields* f, v3 pos, int cell_id, double current_depth)
{
    const int L = m->n_levels;
    const int LP1 = L + 1;
    if (cell_id < 0 || L <= 1 || L > ORC_MAX_VERTICAL_LEVEL_NUM) return fail_state(ORC_ST_BAD_SETUP);
    const int nv = m->n_edges_on_cell[cell_id];
    if (nv <= 0 || nv > ORC_MAX_VERTEX_NUM) return fail_state(ORC_ST_BAD_SETUP);
    if (!is_in_mesh(m, cell_id, pos)) return fail_state(ORC_ST_NOT_IN_CELL);

    int64_t vidx[ORC_MAX_VERTEX_NUM];
    double w[ORC_MAX_VERTEX_NUM];
    cell_weights(m, cell_id, pos, nv, vidx, w);

    double col[ORC_MAX_VERTICAL_LEVEL_NUM];
    if (point_column(m, vidx, w, nv, f->ztop_v, col) != 0) return fail_state(ORC_ST_BAD_COLUMN);

    /* layer search, VK:791-822 (binary, eps 1e-8) */
    const double eps = 1e-8;
    int local_layer = -1;
    if (current_depth > col[0] + eps) {
        local_layer = 1;
    } else if (current_depth < col[L - 1] - eps) {
        local_layer = L - 1;
    } else {
        int lo = 1, hi = L - 1, ans = 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const double top_i = col[mid - 1];
            const double bot_i = col[mid];
            if (current_depth <= top_i + eps && current_depth >= bot_i - eps) { ans = mid; break; }
            if (current_depth > top_i + eps) hi = mid - 1; else lo = mid + 1;
        }
        if (ans < 1) ans = 1;
        if (ans > L - 1) ans = L - 1;
        local_layer = ans;
    }
    if (local_layer < 0) return fail_state(ORC_ST_BAD_COLUMN);

    const double ztop_dn = col[local_layer];
    const double ztop_up = col[local_layer - 1];
    /* x = std::max(dn, std::min(depth, up)), VK:830-831, with the libstdc++ definitions
     * std::min(a,b) = (b<a)?b:a and std::max(a,b) = (a<b)?b:a so NaNs take the same path */
    const double mn = (ztop_up < current_depth) ? ztop_up : current_depth;
    const double x = (ztop_dn < mn) ? mn : ztop_dn;
    const double denom = ztop_up - ztop_dn;
    if (fabs(denom) < 1e-12) return fail_state(ORC_ST_BAD_COLUMN);
    const double t = (x - ztop_dn) / denom;

    v3 vel_dn = calc_velocity(vidx, w, nv, L, local_layer, f->vel_v);
    v3 vel_up = calc_velocity(vidx, w, nv, L, local_layer - 1, f->vel_v);
    if (v3_len(vel_dn) < 1e-12 || v3_len(vel_up) < 1e-12) return fail_state(ORC_ST_ZERO_VELOCITY);
    v3 final_vel = v3_add(v3_mul(vel_up, t), v3_mul(vel_dn, (1.0 - t)));
    if (v3_len(final_vel) < 1e-12) return fail_state(ORC_ST_ZERO_VELOCITY);

    int dn_if = local_layer;
    int up_if = (local_layer > 0) ? (local_layer - 1) : 0;
    if (dn_if >= LP1) dn_if = LP1 - 1;
    if (up_if >= LP1) up_if = LP1 - 1;
    const double w_dn = calc_attribute(vidx, w, nv, LP1, dn_if, f->w_v);
    const double w_up = calc_attribute(vidx, w, nv, LP1, up_if, f->w_v);
    const double vertical_vel = t * w_up + (1.0 - t) * w_dn;

    vel_state s;
    s.h_vel = final_vel; s.v_vel = vertical_vel; s.attr = v3_make(0.0, 0.0, 0.0); s.ok = 1; s.why = 0; s.layer = local_layer;
    return s;
}

/* ------------------------------------------------------------------------------------- */
/* a13: StreamLine, VK:653-1015 (per-particle loop VK:874-1003)                           */
/* ------------------------------------------------------------------------------------- */
/* out_pos/out_vel: [n][each][3], MUST be zero-initialised by the caller (the reference's
 * value-initialised buffers, src/Common/TrajectoryCommon.h:20-25).
 * cell_log: [n][times] or NULL; -1 where the step was not executed.
 * steps_alive: [n] number of steps that were started (alive at step start). */
int orc_streamline(int n_cells, int n_vertices, int max_edges, int L,
                   const double* cell_xyz, const double* vertex_xyz, const int32_t* voc, const int32_t* coc, const int32_t* nedges,
                   const double* ztop_v, const double* vel_v, const double* w_v,
                   int method_rk4, int forward, int64_t deltaT, int64_t duration, int64_t recordT,
                   int64_t n, double* pos_inout, float* depth_inout, const int32_t* cell0,
                   double* out_pos, double* out_vel, int32_t* cell_log, int32_t* steps_alive, int32_t* status,
                   int32_t* final_cell)
{
    if (deltaT == 0 || recordT == 0 || duration == 0) return -1; /* VK:666-669 */
    orc_mesh m = {n_cells, n_vertices, max_edges, L, cell_xyz, vertex_xyz, voc, coc, nedges};
    orc_fields f = {ztop_v, vel_v, w_v, NULL, NULL};
    const int each = (int)(duration / recordT);
    const int times = (int)(duration / deltaT);
    const int dt_sign = forward ? 1 : -1;
    const int delta_t = dt_sign * (int)deltaT;
    const int use_euler = !method_rk4;
    if (each <= 0 || times <= 0) return -2;

    for (int64_t pid = 0; pid < n; ++pid) {
        int run_time = 0;
        int first_loop = 1, first_vel = 1;
        const int64_t base_idx = pid * each;
        int update_points_idx = 0;
        int cell_id = -1;
        int neig[ORC_MAX_CELL_NEIGHBOR_NUM];
        for (int i = 0; i < ORC_MAX_CELL_NEIGHBOR_NUM; ++i) neig[i] = -1;
        int st = ORC_ST_ALIVE;
        int started = 0;
        if (cell_log) for (int t = 0; t < times; ++t) cell_log[pid * times + t] = -1;

        for (int times_i = 0; times_i < times; ++times_i) {
            run_time += abs(delta_t);
            v3 sample = v3_ld(pos_inout, pid);
            const double current_depth = -1.0 * (double)depth_inout[pid];

            if (first_loop) {
                first_loop = 0;
                cell_id = cell0[pid];
                if (cell_id < 0 || cell_id >= n_cells) { st = ORC_ST_BAD_CELL; break; }
                get_cell_neighbors(&m, cell_id, nedges[cell_id], neig);
                v3_st(out_pos, base_idx, sample);
            } else {
                if (cell_id < 0 || cell_id >= n_cells) { st = ORC_ST_BAD_CELL; break; }
                cell_id = relocate(&m, cell_id, neig, sample);
            }
            ++started;
            if (cell_log) cell_log[pid * times + times_i] = cell_id;

            const v3 cur = sample;
            const double r = v3_len(cur);
            v3 rk4_next = cur;
            v3 hvel = v3_make(0.0, 0.0, 0.0);
            double vvel = 0.0;

            if (use_euler) {
                vel_state s = calc_velocity_at_stream(&m, &f, cur, cell_id, current_depth);
                if (!s.ok) { st = s.why; break; }
                hvel = s.h_vel; vvel = s.v_vel;
            } else {
                const double dt = (double)delta_t;
                vel_state s1 = calc_velocity_at_stream(&m, &f, cur, cell_id, current_depth);
                if (!s1.ok) { st = s1.why; break; }
                v3 p2 = advect_on_sphere(cur, s1.h_vel, dt * 0.5);
                vel_state s2 = calc_velocity_at_stream(&m, &f, p2, cell_id, current_depth);
                if (!s2.ok) { st = s2.why; break; }
                v3 p3 = advect_on_sphere(cur, s2.h_vel, dt * 0.5);
                vel_state s3 = calc_velocity_at_stream(&m, &f, p3, cell_id, current_depth);
                if (!s3.ok) { st = s3.why; break; }
                v3 p4 = advect_on_sphere(cur, s3.h_vel, dt);
                vel_state s4 = calc_velocity_at_stream(&m, &f, p4, cell_id, current_depth);
                if (!s4.ok) { st = s4.why; break; }
                /* VK:959-964 */
                hvel = v3_div(v3_add(v3_add(v3_add(s1.h_vel, v3_mul(s2.h_vel, 2.0)), v3_mul(s3.h_vel, 2.0)), s4.h_vel), 6.0);
                vvel = (s1.v_vel + 2.0 * s2.v_vel + 2.0 * s3.v_vel + s4.v_vel) / 6.0;
                v3 x_trial = v3_add(cur, v3_mul(hvel, dt));
                const double x_trial_len = v3_len(x_trial);
                rk4_next = (x_trial_len > 1e-12) ? v3_mul(v3_div(x_trial, x_trial_len), r) : cur;
            }

            v3 new_pos;
            if (use_euler) { /* VK:968-972 */
                v3 axis = rotation_axis(cur, hvel);
                const double speed = v3_len(hvel);
                const double theta_rad = (speed * delta_t) / ((1e-12 < r) ? r : 1e-12);
                new_pos = position_after_rotation(cur, axis, theta_rad);
            } else {
                new_pos = rk4_next;
            }

            /* VK:977-986 */
            const double old_depth = (double)depth_inout[pid];
            double new_depth = old_depth - vvel * (double)delta_t;
            new_depth = (0.0 < new_depth) ? new_depth : 0.0;
            const double r_sum = r + vvel * (double)delta_t;
            const double r_new = (1.0 < r_sum) ? r_sum : 1.0;
            depth_inout[pid] = (float)new_depth;
            const double nlen = v3_len(new_pos);
            if (nlen > 1e-12) new_pos = v3_mul(v3_div(new_pos, nlen), r_new);

            if (first_vel) { first_vel = 0; v3_st(out_vel, base_idx, hvel); } /* VK:988-991 */

            v3_st(pos_inout, pid, new_pos);
            if (recordT > 0 && (run_time % (int)recordT) == 0) { /* VK:994-1001 */
                int64_t write_idx = base_idx + update_points_idx;
                if (write_idx >= base_idx && write_idx < base_idx + each) {
                    v3_st(out_pos, write_idx, new_pos);
                    v3_st(out_vel, write_idx, hvel);
                }
                ++update_points_idx;
            }
        }
        if (steps_alive) steps_alive[pid] = started;
        if (status) status[pid] = st;
        if (final_cell) final_cell[pid] = cell_id;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* a9': calc_velocity_at (pathline), VK:1124-1327                                         */
/* ------------------------------------------------------------------------------------- */
static int linear_layer(const double* col, int L, double d, int* skip_above)
{
    /* VK:1182-1218: skip branches, else first k in 1..L-1 with d <= z[k-1]+eps && d >= z[k]-eps */
    const double eps = 1e-8;
    *skip_above = 0;
    if (d > col[0] + eps) { *skip_above = 1; return 0; }
    if (d < col[L - 1] - eps) return L - 1;
    for (int k = 1; k < L; ++k)
        if (d <= col[k - 1] + eps && d >= col[k] - eps) return k;
    return -1;
}

static vel_state calc_velocity_at_path(const orc_mesh* m, const orc_fields* ff, const orc_fields* fb, int attr_count,
                                       v3 pos, int cell_id, double current_depth, double alpha)
{
    const int L = m->n_levels;
    const int LP1 = L + 1;
    if (cell_id < 0 || L <= 1 || L > ORC_MAX_VERTICAL_LEVEL_NUM) return fail_state(ORC_ST_BAD_SETUP);
    const int nv = m->n_edges_on_cell[cell_id];
    if (nv <= 0 || nv > ORC_MAX_VERTEX_NUM) return fail_state(ORC_ST_BAD_SETUP);
    if (!is_in_mesh(m, cell_id, pos)) return fail_state(ORC_ST_NOT_IN_CELL);

    int64_t vidx[ORC_MAX_VERTEX_NUM];
    double w[ORC_MAX_VERTEX_NUM];
    cell_weights(m, cell_id, pos, nv, vidx, w);

    double zf[ORC_MAX_VERTICAL_LEVEL_NUM], zb[ORC_MAX_VERTICAL_LEVEL_NUM];
    if (point_column(m, vidx, w, nv, ff->ztop_v, zf) != 0) return fail_state(ORC_ST_BAD_COLUMN);
    if (point_column(m, vidx, w, nv, fb->ztop_v, zb) != 0) return fail_state(ORC_ST_BAD_COLUMN);

    int above_f = 0, above_b = 0;
    const int lf = linear_layer(zf, L, current_depth, &above_f);
    const int lb = linear_layer(zb, L, current_depth, &above_b);
    if (above_f || above_b) return fail_state(ORC_ST_ABOVE_SURFACE); /* N2: reference reads ztop[-1] here */
    if (lf < 0 || lb < 0) return fail_state(ORC_ST_BAD_COLUMN);

    const double f_dn = zf[lf], f_up = zf[lf - 1];
    const double b_dn = zb[lb], b_up = zb[lb - 1];
    double mn = (f_up < current_depth) ? f_up : current_depth;
    double x_front = (f_dn < mn) ? mn : f_dn;
    double denom_front = f_up - f_dn;
    if (fabs(denom_front) < 1e-12) return fail_state(ORC_ST_BAD_COLUMN);
    double t_front = (x_front - f_dn) / denom_front;
    mn = (b_up < current_depth) ? b_up : current_depth;
    double x_back = (b_dn < mn) ? mn : b_dn;
    double denom_back = b_up - b_dn;
    if (fabs(denom_back) < 1e-12) return fail_state(ORC_ST_BAD_COLUMN);
    double t_back = (x_back - b_dn) / denom_back;

    v3 vdf = calc_velocity(vidx, w, nv, L, lf, ff->vel_v);
    v3 vuf = calc_velocity(vidx, w, nv, L, lf - 1, ff->vel_v);
    v3 vel_front = v3_add(v3_mul(vuf, t_front), v3_mul(vdf, (1.0 - t_front)));
    v3 vdb = calc_velocity(vidx, w, nv, L, lb, fb->vel_v);
    v3 vub = calc_velocity(vidx, w, nv, L, lb - 1, fb->vel_v);
    v3 vel_back = v3_add(v3_mul(vub, t_back), v3_mul(vdb, (1.0 - t_back)));
    v3 hvel = v3_add(v3_mul(vel_back, alpha), v3_mul(vel_front, (1.0 - alpha)));

    int dn_f = lf, up_f = (lf > 0) ? (lf - 1) : 0, dn_b = lb, up_b = (lb > 0) ? (lb - 1) : 0;
    if (dn_f >= LP1) dn_f = LP1 - 1;
    if (up_f >= LP1) up_f = LP1 - 1;
    if (dn_b >= LP1) dn_b = LP1 - 1;
    if (up_b >= LP1) up_b = LP1 - 1;
    double w_dn_f = calc_attribute(vidx, w, nv, LP1, dn_f, ff->w_v);
    double w_up_f = calc_attribute(vidx, w, nv, LP1, up_f, ff->w_v);
    double w_front = t_front * w_up_f + (1.0 - t_front) * w_dn_f;
    double w_dn_b = calc_attribute(vidx, w, nv, LP1, dn_b, fb->w_v);
    double w_up_b = calc_attribute(vidx, w, nv, LP1, up_b, fb->w_v);
    double w_back = t_back * w_up_b + (1.0 - t_back) * w_dn_b;
    double vvel = alpha * w_back + (1.0 - alpha) * w_front;

    v3 attr = v3_make(0.0, 0.0, 0.0);
    for (int a = 0; a < attr_count && a < 2; ++a) { /* VK:1288-1324 */
        const double* af = a == 0 ? ff->attr0_v : ff->attr1_v;
        const double* ab = a == 0 ? fb->attr0_v : fb->attr1_v;
        double adf = calc_attribute(vidx, w, nv, L, lf, af);
        double auf = calc_attribute(vidx, w, nv, L, lf - 1, af);
        double a_front = t_front * auf + (1.0 - t_front) * adf;
        double adb = calc_attribute(vidx, w, nv, L, lb, ab);
        double aub = calc_attribute(vidx, w, nv, L, lb - 1, ab);
        double a_back = t_back * aub + (1.0 - t_back) * adb;
        double val = alpha * a_back + (1.0 - alpha) * a_front;
        if (a == 0) attr.x = val; else attr.y = val;
    }

    vel_state s;
    s.h_vel = hvel; s.v_vel = vvel; s.attr = attr; s.ok = 1; s.why = 0; s.layer = lf;
    return s;
}

static double clamp01(double v) { return (v < 0.0) ? 0.0 : ((1.0 < v) ? 1.0 : v); } /* std::clamp */

/* a14: PathLine, VK:1017-1496 (per-particle loop VK:1329-1483) */
int orc_pathline(int n_cells, int n_vertices, int max_edges, int L,
                 const double* cell_xyz, const double* vertex_xyz, const int32_t* voc, const int32_t* coc, const int32_t* nedges,
                 const double* ztop_f, const double* vel_f, const double* w_f, const double* a0_f, const double* a1_f,
                 const double* ztop_b, const double* vel_b, const double* w_b, const double* a0_b, const double* a1_b,
                 int attr_count,
                 int method_rk4, int forward, int64_t deltaT, int64_t duration, int64_t recordT,
                 int64_t n, double* pos_inout, float* depth_inout, const int32_t* cell0,
                 double* out_pos, double* out_vel, double* out_attr, int32_t* cell_log, int32_t* steps_alive, int32_t* status,
                 int32_t* final_cell)
{
    if (deltaT == 0 || recordT == 0 || duration == 0) return -1;
    orc_mesh m = {n_cells, n_vertices, max_edges, L, cell_xyz, vertex_xyz, voc, coc, nedges};
    orc_fields ff = {ztop_f, vel_f, w_f, a0_f, a1_f};
    orc_fields fb = {ztop_b, vel_b, w_b, a0_b, a1_b};
    const int each = (int)(duration / recordT);
    const int n_steps = (int)(duration / deltaT);
    const int dt_sign = forward ? 1 : -1;
    const int delta_t = dt_sign * (int)deltaT;
    const int use_euler = !method_rk4;
    const int has_attr = attr_count > 0;
    if (each <= 0 || n_steps <= 0) return -2;

    for (int64_t pid = 0; pid < n; ++pid) {
        int first_loop = 1, first_vel = 1, first_attr = 1;
        const int64_t base_idx = pid * each;
        int update_points_idx = 0;
        int cell_id = -1;
        int neig[ORC_MAX_CELL_NEIGHBOR_NUM];
        for (int i = 0; i < ORC_MAX_CELL_NEIGHBOR_NUM; ++i) neig[i] = -1;
        int st = ORC_ST_ALIVE;
        int started = 0;
        if (cell_log) for (int t = 0; t < n_steps; ++t) cell_log[pid * n_steps + t] = -1;

        for (int step_i = 0; step_i < n_steps; ++step_i) {
            const double alpha = (double)step_i / (double)n_steps;
            v3 sample = v3_ld(pos_inout, pid);
            const double current_depth = -1.0 * (double)depth_inout[pid];
            if (first_loop) {
                first_loop = 0;
                cell_id = cell0[pid];
                if (cell_id < 0 || cell_id >= n_cells) { st = ORC_ST_BAD_CELL; break; }
                get_cell_neighbors(&m, cell_id, nedges[cell_id], neig);
                v3_st(out_pos, base_idx, sample);
            } else {
                if (cell_id < 0 || cell_id >= n_cells) { st = ORC_ST_BAD_CELL; break; }
                cell_id = relocate(&m, cell_id, neig, sample);
            }
            ++started;
            if (cell_log) cell_log[pid * n_steps + step_i] = cell_id;

            const v3 cur = sample;
            const double r = v3_len(cur);
            v3 rk4_next = cur;
            v3 hvel = v3_make(0.0, 0.0, 0.0), attrs = v3_make(0.0, 0.0, 0.0);
            double vvel = 0.0;

            if (use_euler) {
                vel_state s = calc_velocity_at_path(&m, &ff, &fb, attr_count, cur, cell_id, current_depth, alpha);
                if (!s.ok) { st = s.why; break; }
                hvel = s.h_vel; vvel = s.v_vel; attrs = s.attr;
            } else {
                const double dt = (double)delta_t;
                const double dalpha = dt / (double)duration; /* VK:1401 */
                double a1 = alpha;
                vel_state s1 = calc_velocity_at_path(&m, &ff, &fb, attr_count, cur, cell_id, current_depth, a1);
                if (!s1.ok) { st = s1.why; break; }
                v3 p2 = advect_on_sphere(cur, s1.h_vel, dt * 0.5);
                double a2 = clamp01(a1 + 0.5 * dalpha);
                vel_state s2 = calc_velocity_at_path(&m, &ff, &fb, attr_count, p2, cell_id, current_depth, a2);
                if (!s2.ok) { st = s2.why; break; }
                v3 p3 = advect_on_sphere(cur, s2.h_vel, dt * 0.5);
                double a3 = clamp01(a1 + 0.5 * dalpha);
                vel_state s3 = calc_velocity_at_path(&m, &ff, &fb, attr_count, p3, cell_id, current_depth, a3);
                if (!s3.ok) { st = s3.why; break; }
                v3 p4 = advect_on_sphere(cur, s3.h_vel, dt);
                double a4 = clamp01(a1 + dalpha);
                vel_state s4 = calc_velocity_at_path(&m, &ff, &fb, attr_count, p4, cell_id, current_depth, a4);
                if (!s4.ok) { st = s4.why; break; }
                hvel = v3_div(v3_add(v3_add(v3_add(s1.h_vel, v3_mul(s2.h_vel, 2.0)), v3_mul(s3.h_vel, 2.0)), s4.h_vel), 6.0);
                attrs = v3_div(v3_add(v3_add(v3_add(s1.attr, v3_mul(s2.attr, 2.0)), v3_mul(s3.attr, 2.0)), s4.attr), 6.0);
                vvel = (s1.v_vel + 2.0 * s2.v_vel + 2.0 * s3.v_vel + s4.v_vel) / 6.0;
                v3 x_trial = v3_add(cur, v3_mul(hvel, dt));
                double x_trial_len = v3_len(x_trial);
                rk4_next = (x_trial_len > 1e-12) ? v3_mul(v3_div(x_trial, x_trial_len), r) : cur;
            }

            v3 new_pos;
            if (use_euler) {
                v3 axis = rotation_axis(cur, hvel);
                double speed = v3_len(hvel);
                double theta_rad = (speed * delta_t) / ((1e-12 < r) ? r : 1e-12);
                new_pos = position_after_rotation(cur, axis, theta_rad);
            } else {
                new_pos = rk4_next;
            }

            if (first_vel) { first_vel = 0; v3_st(out_vel, base_idx, hvel); }
            if (first_attr && has_attr) { first_attr = 0; v3_st(out_attr, base_idx, attrs); }

            const double old_depth = (double)depth_inout[pid];
            double new_depth = old_depth - vvel * (double)delta_t;
            new_depth = (0.0 < new_depth) ? new_depth : 0.0;
            const double r_sum = r + vvel * (double)delta_t;
            double r_new = (1.0 < r_sum) ? r_sum : 1.0;
            depth_inout[pid] = (float)new_depth;
            const double nlen = v3_len(new_pos);
            if (nlen > 1e-12) new_pos = v3_mul(v3_div(new_pos, nlen), r_new);
            v3_st(pos_inout, pid, new_pos);

            const int record_interval = (int)(recordT / deltaT); /* VK:1470-1481 */
            if (record_interval > 0 && ((step_i + 1) % record_interval) == 0) {
                int64_t write_idx = base_idx + update_points_idx;
                if (write_idx >= base_idx && write_idx < base_idx + each) {
                    v3_st(out_pos, write_idx, new_pos);
                    v3_st(out_vel, write_idx, hvel);
                    if (has_attr) v3_st(out_attr, write_idx, attrs);
                }
                ++update_points_idx;
            }
        }
        if (steps_alive) steps_alive[pid] = started;
        if (status) status[pid] = st;
        if (final_cell) final_cell[pid] = cell_id;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* a15: VisualizeFixedDepth, VK:238-471                                                   */
/* ------------------------------------------------------------------------------------- */

/* pixel (i = row, j = col) -> sample position; GeoConverter.hpp:9-33 + :107-125 */
void orc_pixel_position(int width, int height, double minLat, double maxLat, double minLon, double maxLon,
                        int i, int j, double* out3)
{
    double lat = maxLat - ((double)i / (double)height * (maxLat - minLat));
    double lon = ((double)j / (double)width * (maxLon - minLon)) + minLon;
    lat = lat * (M_PI / 180.0);
    lon = lon * (M_PI / 180.0);
    const double r = 6371010.0f;
    double costheta = cos(lat), cosphi = cos(lon);
    double sintheta = sin(lat), sinphi = sin(lon);
    out3[0] = r * costheta * cosphi;
    out3[1] = r * costheta * sinphi;
    out3[2] = r * sintheta;
}

static void set_pixel(double* img, int w, int h, int i, int j, double a, double b, double c)
{
    if (i < 0 || i >= h || j < 0 || j >= w) return;
    int64_t index = ((int64_t)i * w + j) * 4;
    img[index + 0] = a; img[index + 1] = b; img[index + 2] = c; img[index + 3] = 1.0;
}

/* img0 (+img1 when n_attr_total > 1, the reference's `mDoubleAttributes.size() > 1`):
 * [height][width][4], caller zero-initialises.  pixel_cell: [height*width] (cell ids as
 * located; may be passed in pre-computed when `have_cells` != 0). */
int orc_remap_fixed_depth(int n_cells, int n_vertices, int max_edges, int L,
                          const double* cell_xyz, const double* vertex_xyz, const int32_t* voc, const int32_t* coc, const int32_t* nedges,
                          const double* ztop_v, const double* vel_v, const double* a0_v, const double* a1_v, int attr_count,
                          int width, int height, double minLat, double maxLat, double minLon, double maxLon, double fixed_depth,
                          double* img0, double* img1, int32_t* pixel_cell, int have_cells)
{
    orc_mesh m = {n_cells, n_vertices, max_edges, L, cell_xyz, vertex_xyz, voc, coc, nedges};
    const double DEPTH = -fixed_depth; /* VK:251 */
    const double dnan = NAN;
    const int bAttr = (attr_count > 1) && img1 != NULL; /* VK:259-267 */
    orc_bins* bins = have_cells ? NULL : bins_build(n_cells, cell_xyz);

    for (int64_t gid = 0; gid < (int64_t)width * height; ++gid) {
        const int ih = (int)(gid / width);
        const int jw = (int)(gid % width);
        double pp[3];
        orc_pixel_position(width, height, minLat, maxLat, minLon, maxLon, ih, jw, pp);
        v3 pos = v3_make(pp[0], pp[1], pp[2]);
        int cell_id = have_cells ? pixel_cell[gid] : bins_nearest(bins, n_cells, cell_xyz, pp);
        if (!have_cells && pixel_cell) pixel_cell[gid] = cell_id;

#define ORC_NAN_PIXEL() do { set_pixel(img0, width, height, ih, jw, dnan, dnan, dnan); \
                             if (bAttr) set_pixel(img1, width, height, ih, jw, dnan, dnan, dnan); } while (0)
        if (cell_id < 0 || cell_id >= n_cells) { ORC_NAN_PIXEL(); continue; }
        const int nv = nedges[cell_id];
        if (!is_in_mesh(&m, cell_id, pos)) { ORC_NAN_PIXEL(); continue; }
        if (nv > ORC_MAX_VERTEX_NUM) { ORC_NAN_PIXEL(); continue; }
        int64_t vidx[ORC_MAX_VERTEX_NUM];
        double w[ORC_MAX_VERTEX_NUM];
        cell_weights(&m, cell_id, pos, nv, vidx, w);
        if (L <= 0 || L > ORC_MAX_VERTICAL_LEVEL_NUM) { ORC_NAN_PIXEL(); continue; }

        double col[ORC_MAX_VERTICAL_LEVEL_NUM];
        for (int k = 0; k < L; ++k) { /* VK:346-354 */
            double acc = 0.0;
            for (int v = 0; v < nv; ++v) acc += w[v] * ztop_v[vidx[v] * L + k];
            col[k] = acc;
        }
        for (int k = 1; k < L; ++k)
            if (col[k] > col[k - 1]) col[k] = col[k - 1] - 1e-9;

        double z_surf = col[0], z_bot = col[L - 1];
        if (z_surf < z_bot) { double t = z_surf; z_surf = z_bot; z_bot = t; }
        double ad = 1e-8 * fabs(z_surf - z_bot);
        double epsd = (1e-6 < ad) ? ad : 1e-6; /* std::max(1e-6, ad) */
        if (!(DEPTH <= z_surf + epsd && DEPTH >= z_bot - epsd)) { ORC_NAN_PIXEL(); continue; }

        int local_layer = -1;
        for (int k = 1; k < L; ++k) { /* VK:378-391 */
            double topI = col[k - 1], botI = col[k];
            if (topI < botI) { double t = topI; topI = botI; botI = t; }
            if (DEPTH <= topI + 1e-8 && DEPTH >= botI - 1e-8) { local_layer = k; break; }
        }
        if (DEPTH <= col[0]) local_layer = 0; /* VK:392-394: the override */
        if (local_layer < 0) { ORC_NAN_PIXEL(); continue; }

        double topI = col[(local_layer - 1 > 0) ? local_layer - 1 : 0];
        double botI = col[local_layer];
        if (topI < botI) { double t = topI; topI = botI; botI = t; }
        double denom = topI - botI;
        double tparam = (denom > 1e-12) ? (DEPTH - botI) / denom : 0.5;

        const int vel_levels = L;
        int j = local_layer - 1;
        if (j < 0) j = 0;
        if (j > vel_levels - 1) j = vel_levels - 1;
        int j_bot = (j + 1 < vel_levels - 1) ? j + 1 : vel_levels - 1;
        int j_top = j;
        v3 v_top = calc_velocity(vidx, w, nv, vel_levels, j_top, vel_v);
        v3 v_bot = calc_velocity(vidx, w, nv, vel_levels, j_bot, vel_v);
        double mtop = v3_len(v_top), mbot = v3_len(v_bot);
        v3 final_vel;
        if (mtop < 1e-12 && mbot < 1e-12) final_vel = v3_make(0.0, 0.0, 0.0);
        else if (mtop < 1e-12) final_vel = v_bot;
        else if (mbot < 1e-12) final_vel = v_top;
        else final_vel = v3_add(v3_mul(v_bot, (1.0 - tparam)), v3_mul(v_top, tparam));

        double u_east, v_north;
        xyz_to_enu(pos, final_vel, &u_east, &v_north);
        double spd = sqrt(u_east * u_east + v_north * v_north);

        double a0 = 0.0, a1 = 0.0;
        if (bAttr) { /* VK:443-464 */
            int aj = local_layer - 1;
            if (aj < 0) aj = 0;
            if (aj > L - 1) aj = L - 1;
            if (attr_count >= 1) a0 = calc_attribute(vidx, w, nv, L, aj, a0_v);
            if (attr_count >= 2) a1 = calc_attribute(vidx, w, nv, L, aj, a1_v);
        }
        set_pixel(img0, width, height, ih, jw, u_east, v_north, spd);
        if (bAttr) set_pixel(img1, width, height, ih, jw, a0, a1, 0.0);
#undef ORC_NAN_PIXEL
    }
    bins_free(bins);
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* 8f-3: VisualizeFixedLayer (VK:141-236) and VisualizeFixedLatitude (VK:473-651)            */
/* ------------------------------------------------------------------------------------- */
int orc_remap_fixed_layer(int n_cells, int n_vertices, int max_edges, int L,
                          const double* cell_xyz, const double* vertex_xyz, const int32_t* voc, const int32_t* coc, const int32_t* nedges,
                          const double* vel_v, int width, int height, double minLat, double maxLat, double minLon, double maxLon,
                          int fixed_layer_in, double* img, int32_t* pixel_cell)
{
    orc_mesh m = {n_cells, n_vertices, max_edges, L, cell_xyz, vertex_xyz, voc, coc, nedges};
    int fixed_layer = fixed_layer_in; /* ClampLayer, VK:14-26 */
    if (L <= 0) fixed_layer = 0;
    else if (fixed_layer < 0) fixed_layer = 0;
    else if (fixed_layer >= L) fixed_layer = L - 1;
    orc_bins* bins = bins_build(n_cells, cell_xyz);
    for (int64_t gid = 0; gid < (int64_t)width * height; ++gid) {
        const int ih = (int)(gid / width), jw = (int)(gid % width);
        double pp[3];
        orc_pixel_position(width, height, minLat, maxLat, minLon, maxLon, ih, jw, pp);
        v3 pos = v3_make(pp[0], pp[1], pp[2]);
        const int cell_id = bins_nearest(bins, n_cells, cell_xyz, pp);
        if (pixel_cell) pixel_cell[gid] = cell_id;
        int ok = !(cell_id < 0 || cell_id >= n_cells);
        int nv = 0;
        if (ok) { nv = nedges[cell_id]; ok = !(nv <= 0 || nv > ORC_MAX_VERTEX_NUM); }
        if (ok) ok = is_in_mesh(&m, cell_id, pos);
        if (!ok) { set_pixel(img, width, height, ih, jw, NAN, NAN, NAN); continue; }
        int64_t vidx[ORC_MAX_VERTEX_NUM];
        double w[ORC_MAX_VERTEX_NUM];
        cell_weights(&m, cell_id, pos, nv, vidx, w);
        v3 vel = calc_velocity(vidx, w, nv, L, fixed_layer, vel_v);
        double zon, mer;
        xyz_to_enu(pos, vel, &zon, &mer);
        set_pixel(img, width, height, ih, jw, zon, mer, 0.0);
    }
    bins_free(bins);
    return 0;
}

/* MPASOField::isOnOcean, src/Core/MPASOField.cpp:36-81: returns is_land */
static int is_land(const orc_mesh* m, int cell_id, v3 position, int nv)
{
    double first = 0.0;
    for (int k = 0; k < nv; ++k) {
        int64_t A_idx = (int64_t)m->vertices_on_cell[(int64_t)cell_id * m->max_edges + k] - 1;
        int64_t B_idx = (int64_t)m->vertices_on_cell[(int64_t)cell_id * m->max_edges + ((k + 1) % nv)] - 1;
        v3 A = v3_ld(m->vertex_xyz, A_idx), B = v3_ld(m->vertex_xyz, B_idx);
        v3 O = v3_make(0.0, 0.0, 0.0);
        v3 AO = v3_sub(O, A), BO = v3_sub(O, B), A_point = v3_sub(position, A);
        v3 surface_normal = v3_cross(AO, BO);
        double direction = v3_dot(surface_normal, A_point);
        double sign = (direction > 0) ? 1.0 : -1.0;
        if (k == 0) first = sign;
        else if (sign != first) return 1;
    }
    return 0;
}

int orc_regrid_fixed_latitude(int n_cells, int n_vertices, int max_edges, int L,
                              const double* cell_xyz, const double* vertex_xyz, const int32_t* voc, const int32_t* coc, const int32_t* nedges,
                              const double* ztop_v, const double* vel_v, int width, int height, double minLon, double maxLon,
                              double fixed_lat, double minDepth, double maxDepth, double* img, int32_t* pixel_cell)
{
    orc_mesh m = {n_cells, n_vertices, max_edges, L, cell_xyz, vertex_xyz, voc, coc, nedges};
    const int nVert = L;
    const double i_step = (height > 1) ? (maxDepth - minDepth) / (height - 1) : 0.0;
    const double j_step = (width > 1) ? (maxLon - minLon) / (width - 1) : 0.0;
    orc_bins* bins = bins_build(n_cells, cell_xyz);
    for (int64_t gid = 0; gid < (int64_t)width * height; ++gid) {
        const int ih = (int)(gid / width), jw = (int)(gid % width);
        const double depth_plot = minDepth + ih * i_step;
        const double DEPTH = -fabs(depth_plot);
        const double lon = minLon + jw * j_step;
        /* convertRadianLatLonToXYZ */
        const double theta = fixed_lat * (M_PI / 180.0), phi = lon * (M_PI / 180.0);
        const double r = 6371010.0f;
        double costheta = cos(theta), cosphi = cos(phi), sintheta = sin(theta), sinphi = sin(phi);
        double pp[3] = {r * costheta * cosphi, r * costheta * sinphi, r * sintheta};
        v3 position = v3_make(pp[0], pp[1], pp[2]);
        const int cell_id = bins_nearest(bins, n_cells, cell_xyz, pp);
        if (pixel_cell) pixel_cell[gid] = cell_id;
#define ORC_NANPX() do { set_pixel(img, width, height, ih, jw, NAN, NAN, NAN); } while (0)
        if (cell_id < 0 || cell_id >= n_cells) { ORC_NANPX(); continue; }
        const int nv = nedges[cell_id];
        if (nv <= 0 || nv > ORC_MAX_VERTEX_NUM) { ORC_NANPX(); continue; }
        if (is_land(&m, cell_id, position, nv)) { ORC_NANPX(); continue; }
        int64_t vidx[ORC_MAX_VERTEX_NUM];
        double w[ORC_MAX_VERTEX_NUM];
        cell_weights(&m, cell_id, position, nv, vidx, w);
        double col[ORC_MAX_VERTICAL_LEVEL_NUM];
        for (int k = 0; k < nVert; ++k) {
            double z_acc = 0.0;
            for (int v = 0; v < nv; ++v) z_acc += w[v] * ztop_v[vidx[v] * L + k];
            col[k] = z_acc;
        }
        for (int k = 1; k < nVert; ++k)
            if (col[k] > col[k - 1]) col[k] = col[k - 1] - 1e-9;
        int layer = -1;
        const double EPSILON = 1e-6;
        if (DEPTH > col[0] + EPSILON || DEPTH < col[nVert - 1] - EPSILON) { ORC_NANPX(); continue; }
        for (int k = 1; k < nVert; ++k) {
            double z_up = col[k - 1], z_dn = col[k];
            if (z_up < z_dn) { double t = z_up; z_up = z_dn; z_dn = t; }
            if (DEPTH <= z_up + EPSILON && DEPTH >= z_dn - EPSILON) { layer = k; break; }
        }
        if (layer == -1) { ORC_NANPX(); continue; }
        double dn = col[layer], up = col[layer - 1];
        if (up < dn) { double t = up; up = dn; dn = t; }
        const double denom = up - dn;
        if (fabs(denom) < 1e-30) { ORC_NANPX(); continue; }
        const double t = (DEPTH - dn) / denom;
        v3 vel_up = calc_velocity(vidx, w, nv, L, layer - 1, vel_v);
        v3 vel_dn = calc_velocity(vidx, w, nv, L, layer, vel_v);
        v3 final_vel = v3_add(v3_mul(vel_dn, (1.0 - t)), v3_mul(vel_up, t));
        double zon, mer;
        xyz_to_enu(position, final_vel, &zon, &mer);
        set_pixel(img, width, height, ih, jw, zon, mer, 0.0);
#undef ORC_NANPX
    }
    bins_free(bins);
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* a16: line assembly + NaN trimming, src/Common/TrajectoryCommon.h:43-190                */
/* ------------------------------------------------------------------------------------- */
/* raw_pos/raw_vel: [n][each][3] -> lines: points/velocity [n][each+1][3], last [n][3].
 * points = [seed, rec_0..rec_each-1]; velocity = [vel_0..vel_each-1, 0] (R7);
 * temperature/salinity (pathline) = velocity.x / velocity.y of the same slot (R7), padded 0. */
int orc_finalize_lines(int64_t n, int each, const double* seeds, const double* raw_pos, const double* raw_vel, int with_attrs,
                       double* points, double* velocity, double* temperature, double* salinity, double* last)
{
    const int per = each + 1;
    for (int64_t i = 0; i < n; ++i) {
        double* P = points + i * per * 3;
        double* V = velocity + i * per * 3;
        memcpy(P, seeds + 3 * i, 3 * sizeof(double));
        for (int k = 0; k < each; ++k) {
            memcpy(P + 3 * (k + 1), raw_pos + (i * each + k) * 3, 3 * sizeof(double));
            memcpy(V + 3 * k, raw_vel + (i * each + k) * 3, 3 * sizeof(double));
            if (with_attrs) {
                temperature[i * per + k] = raw_vel[(i * each + k) * 3 + 0];
                salinity[i * per + k] = raw_vel[(i * each + k) * 3 + 1];
            }
        }
        V[3 * each] = V[3 * each + 1] = V[3 * each + 2] = 0.0; /* resize(original_len, 0) */
        if (temperature) { if (!with_attrs) for (int k = 0; k < each; ++k) temperature[i * per + k] = 0.0; temperature[i * per + each] = 0.0; }
        if (salinity) { if (!with_attrs) for (int k = 0; k < each; ++k) salinity[i * per + k] = 0.0; salinity[i * per + each] = 0.0; }

        /* RemoveNaNTrajectoriesAndReindex, :57-129 */
        int k = 0;
        for (; k < per; ++k)
            if (!(isfinite(P[3 * k]) && isfinite(P[3 * k + 1]) && isfinite(P[3 * k + 2]))) break;
        if (k == 0) {
            double first_temp = temperature ? temperature[i * per] : 0.0;
            double first_sal = salinity ? salinity[i * per] : 0.0;
            for (int j = 0; j < per; ++j) {
                memcpy(P + 3 * j, P, 3 * sizeof(double));
                V[3 * j] = V[3 * j + 1] = V[3 * j + 2] = 0.0;
                if (temperature) temperature[i * per + j] = first_temp;
                if (salinity) salinity[i * per + j] = first_sal;
            }
        } else if (k < per) {
            double last_temp = temperature ? temperature[i * per + k - 1] : 0.0;
            double last_sal = salinity ? salinity[i * per + k - 1] : 0.0;
            V[3 * (k - 1)] = V[3 * (k - 1) + 1] = V[3 * (k - 1) + 2] = 0.0;
            for (int j = k; j < per; ++j) {
                memcpy(P + 3 * j, P + 3 * (k - 1), 3 * sizeof(double));
                V[3 * j] = V[3 * j + 1] = V[3 * j + 2] = 0.0;
                if (temperature) temperature[i * per + j] = last_temp;
                if (salinity) salinity[i * per + j] = last_sal;
            }
        }
        if (last) memcpy(last + 3 * i, P + 3 * (per - 1), 3 * sizeof(double));
    }
    return 0;
}

/* small exported probes used by unit tests */
void orc_wachspress(const double* p3, const double* poly, int nv, double* w)
{
    v3 pv[ORC_MAX_VERTEX_NUM];
    for (int i = 0; i < nv; ++i) pv[i] = v3_ld(poly, i);
    wachspress(v3_make(p3[0], p3[1], p3[2]), pv, w, nv);
}

void orc_advect_on_sphere(const double* pos, const double* vel, double dt, double* out)
{
    v3 r = advect_on_sphere(v3_make(pos[0], pos[1], pos[2]), v3_make(vel[0], vel[1], vel[2]), dt);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
