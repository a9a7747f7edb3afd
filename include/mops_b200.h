/*
 * mops_b200.h -- C ABI of the B200-native particle-advection / remap engine.
 *
 * This is the drop-in boundary for the hot path of YosefQiu/MOPS (reference paths below
 * are relative to the reference root).  The innermost stable seam of the reference is
 *
 *   MOPS::Factory::StreamLine / PathLine / VisualizeFixedDepth     src/Common/MOPSFactory.h:9-40
 *   MOPS::Factory::CalcCellVertexZtop / ...Velocity / ...ToVertex  src/Common/MOPSFactory.h:42-106
 *   MPASOField::calcInWhichCells / MPASOGrid::searchKDT            src/Core/MPASOField.cpp:23-34
 *
 * reached from MOPSApp::runStreamLine / runPathLine / runRemapping / addSol
 * (src/Core/MOPSApp.cpp:77-137, 171-196, 231-290).  Every entry point here replaces one of
 * those calls; INTEGRATION.md shows the binding a maintainer adds on the reference side.
 * The C++ drop-in of include/api/MOPS.h that ships in this repo (include/api/MOPS.h) is a
 * thin host layer over exactly these functions.
 *
 * Conventions
 *  - plain pointers and sizes, no C++ / torch types; every function returns 0 on success
 *    or a negative MOPS_E_* code, and mops_last_error(ctx) gives the message.  Nothing
 *    throws across the boundary.  There is NO CPU fallback: without a CUDA device (or if
 *    the sm_100a kernels cannot be loaded) mops_create fails.
 *  - one context = one GPU = one caller thread at a time (the reference API is not
 *    re-entrant either); the one exception is mops_set_snapshot_async for a
 *    slot no running call uses, which may come from a second host thread.  Multi-GPU: the
 *    mops_dist_* (one process per GPU) and mops_multi_* (one process, N GPUs) sections below.
 *  - connectivity is passed exactly as MPAS files / MPASOGrid hold it: int32, 1-based,
 *    0-padded rows of width maxEdges (src/IO/MPASOReader.cpp:147-153).  Cell ids that
 *    cross this boundary in either direction are 0-based indices into the caller's cell
 *    arrays (what searchKDT returns); the engine's internal renumbering is invisible.
 *  - vec3 arrays are [n][3] doubles (the 24-byte layout of the reference's vec3).
 *  - `mem` selects whether particle / image buffers are HOST or DEVICE pointers.  Mesh inputs are
 *    host pointers; snapshot fields are host pointers (pinned memory makes their upload async) or
 *    device pointers (see mops_side_wait_event).
 */
#ifndef MOPS_B200_H
#define MOPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOPS_B200_ABI_VERSION 2

typedef struct mops_ctx mops_ctx;

enum {
    MOPS_OK = 0,
    MOPS_E_INVALID = -1,  /* bad argument / call order (reference: Error(...) + empty result) */
    MOPS_E_CUDA = -2,     /* CUDA runtime error; message in mops_last_error                   */
    MOPS_E_NODEVICE = -3, /* no usable sm_100-class device: there is no CPU fallback          */
    MOPS_E_NOMEM = -4,
    MOPS_E_STATE = -5     /* mesh / snapshot not set                                          */
};

enum { MOPS_MEM_HOST = 0, MOPS_MEM_DEVICE = 1 };

/* CalcMethodType / CalcDirection of the reference, src/Core/MPASOVisualizer.h:16-17 */
enum { MOPS_METHOD_RK4 = 0, MOPS_METHOD_EULER = 1 };
enum { MOPS_DIR_FORWARD = 0, MOPS_DIR_BACKWARD = 1 };
/* MOPS_SEM_REFERENCE: the reference's semantics, bit for bit -- every RK stage is evaluated in the
 * start-of-step cell and a stage point outside it stops the particle for good (VK:753-756, 930-957).
 * MOPS_SEM_WALK: NOT a parity mode -- each stage point is evaluated in the cell that contains it
 * (cellsOnCell walk to the nearest centre), so particles cross cells under RK4 as they do under
 * Euler; validated against the analytic solid-body solution, identical to the reference until the
 * step where the reference would stop the particle. */
enum { MOPS_SEM_REFERENCE = 0, MOPS_SEM_WALK = 1 };

/* why a particle stopped (the reference's kernels just `return`) */
enum {
    MOPS_ST_ALIVE = 0,
    MOPS_ST_BAD_CELL = 1,      /* start cell out of range            VK:895,903        */
    MOPS_ST_NOT_IN_CELL = 2,   /* IsInMesh false at an RK stage      VK:753-756        */
    MOPS_ST_BAD_COLUMN = 3,    /* |z_up - z_dn| < 1e-12 ...          VK:833            */
    MOPS_ST_ZERO_VELOCITY = 4, /* |v| < 1e-12 (streamline only)      VK:845-852        */
    MOPS_ST_ABOVE_SURFACE = 5, /* pathline above-surface branch (reads ztop[-1] in the reference; not replicated) */
    MOPS_ST_BAD_SETUP = 6
};

#define MOPS_MAX_SNAPSHOT_SLOTS 4
#define MOPS_MAX_ATTRS 2 /* the reference's kernels use the first two scalar attributes (R11) */

/* ---- lifetime --------------------------------------------------------------------- */
/* replaces MOPS_Init("gpu") device selection, src/Core/MOPSApp.cpp:34-63 */
int mops_create(mops_ctx** out, int device_ordinal);
void mops_destroy(mops_ctx* ctx);
const char* mops_last_error(const mops_ctx* ctx);
int mops_abi_version(void);
/* pinned host staging for snapshot double-buffering (cudaHostAlloc / cudaFreeHost) */
int mops_host_alloc(void** out, size_t bytes);
int mops_host_free(void* p);
int mops_synchronize(mops_ctx* ctx);
/* run the context's kernels / copies on a caller-owned cudaStream_t (e.g. torch's current
 * stream, so that the caller's own events and NCCL calls order with them); NULL restores the
 * context's private stream.  Snapshot uploads keep using the private side stream. */
int mops_set_stream(mops_ctx* ctx, void* cuda_stream);
/* CUDA-event timing marks on the context's stream: record mark `idx` (0..7) now; elapsed
 * milliseconds between two recorded marks (synchronises on the later one). */
int mops_mark(mops_ctx* ctx, int32_t idx);
int mops_elapsed_ms(mops_ctx* ctx, int32_t idx_from, int32_t idx_to, double* ms_out);

/* ---- mesh: replaces MOPSApp::addGrid + the per-call mesh H2D of the reference's CUDA
 *      wrappers (src/Core/MOPSApp.cpp:65-75; src/GPU/CUDA/Kernel/MPASOVisualizerKernels.cu:1369-1383).
 *      Uploaded once, renumbered along a Morton curve and kept resident in HBM. ------- */
int mops_set_mesh(mops_ctx* ctx, int32_t n_cells, int32_t n_vertices, int32_t max_edges,
                  const double* cell_xyz,            /* [n_cells][3]                          */
                  const double* vertex_xyz,          /* [n_vertices][3]                       */
                  const int32_t* vertices_on_cell,   /* [n_cells][max_edges] 1-based, 0 pad   */
                  const int32_t* cells_on_cell,      /* [n_cells][max_edges] 1-based, 0 pad   */
                  const int32_t* cells_on_vertex,    /* [n_vertices][3] 1-based               */
                  const int32_t* n_edges_on_cell);   /* [n_cells]                             */

/* ---- snapshot: replaces MOPSApp::addSol's preprocessing chain (calcCellCenterZtop,
 *      calcCellVertexZtop, calcCellCenterVelocityByZM, calcCellVertexVelocity,
 *      calcCellVertexVertVelocity, calcCellCenterToVertex; src/Core/MOPSApp.cpp:100-130),
 *      run on the device.  Inputs are cell-major [n_cells][n_levels] host arrays as
 *      MPASOSolution holds them; vert_vel_top is [n_cells][n_levels+1] or NULL (zeros).
 *      attrs: n_attr cell-major scalar fields in the reference's std::map (alphabetical)
 *      order; n_attr_total is mDoubleAttributes.size() (gates the attribute image /
 *      pathline attributes exactly as `size() > 1` does, VK:259-267, 1093-1104).
 *      The async form enqueues upload + preprocessing on the context's side stream and
 *      returns; the next call that uses `slot` waits on it (double-buffered pathlines). */
int mops_set_snapshot(mops_ctx* ctx, int32_t slot, int32_t n_levels,
                      const double* zonal, const double* meridional, const double* layer_thickness,
                      const double* bottom_depth, const double* vert_vel_top,
                      int32_t n_attr, const double* const* attrs, int32_t n_attr_total);
int mops_set_snapshot_async(mops_ctx* ctx, int32_t slot, int32_t n_levels,
                            const double* zonal, const double* meridional, const double* layer_thickness,
                            const double* bottom_depth, const double* vert_vel_top,
                            int32_t n_attr, const double* const* attrs, int32_t n_attr_total);
int mops_snapshot_wait(mops_ctx* ctx, int32_t slot);
/* The field pointers of the two calls above may also be DEVICE pointers (a snapshot assembled on the GPU, e.g. every rank
 * of a multi-GPU job uploads 1/N of the cell-major fields and an all-gather over NVLink completes them; the reference has
 * no counterpart, it re-uploads everything per call, src/GPU/CUDA/Kernel/MPASOVisualizerKernels.cu:2241-2272).  The
 * copies run on the context's side stream: make that stream wait for the producer of the device buffers with
 * mops_side_wait_event (a cudaEvent_t recorded on the producing stream) before mops_set_snapshot_async. */
int mops_side_wait_event(mops_ctx* ctx, void* cuda_event);
/* copy the prepared vertex-major arrays back in the caller's vertex order (parity tests):
 * ztop [nV][L], vel [nV][L][3], vertvel [nV][L+1] (level L is not kept by the engine and
 * reads back as 0 -- no kernel of the path uses it), attr [nV][L].  Any may be NULL.     */
int mops_get_prepared(mops_ctx* ctx, int32_t slot, double* ztop_vertex, double* vel_vertex,
                      double* vertvel_vertex, double* attr0_vertex, double* attr1_vertex);

/* ---- point location: replaces MPASOField::calcInWhichCells (host, serial nanoflann 1-NN;
 *      src/Core/MPASOField.cpp:23-34, src/Core/MPASOGrid.cpp:287-313).  Exact nearest cell
 *      centre by a cooperative cellsOnCell walk started from a lat/lon bucket table.      */
int mops_locate(mops_ctx* ctx, int32_t mem, int64_t n, const double* xyz, int32_t* cell_out);

/* ---- trajectories ----------------------------------------------------------------- */
typedef struct mops_traj_cfg {
    int32_t method;         /* MOPS_METHOD_*   (TrajectorySettings::methodType)              */
    int32_t direction;      /* MOPS_DIR_*      (TrajectorySettings::directionType)           */
    int64_t delta_t;        /* seconds         (TrajectorySettings::deltaT)                  */
    int64_t duration;       /* seconds         (TrajectorySettings::simulationDuration)      */
    int64_t record_t;       /* seconds         (TrajectorySettings::recordT)                 */
    int32_t mem;            /* MOPS_MEM_*: space of all particle / output pointers           */
    int32_t sort_particles; /* 1: process particles in Morton-cell order (results unchanged) */
    int32_t count_near_edge;/* 1: track, per particle, the smallest angular distance [rad] between any
                               evaluated point and an edge of the cell it was evaluated in; particles
                               closer than 1e-12 rad are counted in mops_traj_stats (the band in which
                               a differently-rounded sin/cos may legitimately flip a cell decision) */
    int32_t semantics;      /* MOPS_SEM_REFERENCE (default) or MOPS_SEM_WALK                        */
    int32_t reserved[2];
} mops_traj_cfg;

typedef struct mops_traj_io {
    int64_t n;              /* particles                                                      */
    double* xyz;            /* [n][3]  in: seeds, out: end points (stable_points)             */
    float* depth;           /* [n]     in/out per-particle depth, float as in the reference (R3) */
    const int32_t* cell0;   /* [n] start cells or NULL (then located on the device)           */
    double* out_pos;        /* [n][each][3] recorded positions, each = duration / record_t    */
    double* out_vel;        /* [n][each][3] recorded velocities                               */
    double* out_attr;       /* [n][each][3] pathline attributes (x,y used) or NULL; written only when the snapshots
                               carry attributes (n_attr_total > 1, VK:1093-1104), otherwise left untouched       */
    int32_t* out_cell_log;  /* [n][steps] cell of every step (-1 = not executed) or NULL      */
    int32_t* out_status;    /* [n] MOPS_ST_* or NULL                                          */
    int32_t* out_steps;     /* [n] steps started (alive at step start) or NULL                */
    int32_t* out_cell;      /* [n] last cell or NULL                                          */
    double* out_min_edge;   /* [n] smallest edge distance [rad] seen (count_near_edge) or NULL        */
} mops_traj_io;

typedef struct mops_traj_stats {
    int64_t particle_steps; /* sum of steps started over all particles                        */
    int64_t alive_at_end;
    double kernel_ms;       /* CUDA-event time of the advection kernel(s)                     */
    double locate_ms;
    double total_ms;        /* whole call on the stream (copies included)                     */
    int32_t launches;       /* kernels of this library launched by the call                   */
    int32_t reserved;
    int64_t near_edge_particles; /* particles that came within 1e-12 rad of a cell edge (count_near_edge) */
    int64_t above_surface_particles; /* pathline: particles stopped with MOPS_ST_ABOVE_SURFACE -- the depth lies above the
                                        interpolated sea surface (negative zTop[0]) and the reference would read
                                        ztop[-1] there (VK:1225, undefined behaviour); they keep their last position */
} mops_traj_stats;

/* Both calls integrate in launches of 40 steps with the particles that stopped compacted away between launches, or as
 * one launch when the previous call's particles hardly ever stopped (the engine adapts; environment
 * MOPS_SEGMENT_STEPS=<n> fixes the segment length, 0 = always one launch) -- results are identical either way.  The live
 * count stays on the device, so MOPS_MEM_DEVICE calls with stats = NULL are asynchronous on the context's stream whatever their
 * length.  MOPS_MEM_HOST calls stage through device scratch and complete before they return; their asynchronous
 * form is submit / wait below. */
/* replaces MOPS::Factory::StreamLine (src/Common/MOPSFactory.h:28-33 -> VK:653-1015) */
int mops_streamline(mops_ctx* ctx, const mops_traj_cfg* cfg, int32_t slot, const mops_traj_io* io,
                    mops_traj_stats* stats);
/* replaces MOPS::Factory::PathLine (src/Common/MOPSFactory.h:35-40 -> VK:1017-1496) */
int mops_pathline(mops_ctx* ctx, const mops_traj_cfg* cfg, int32_t front_slot, int32_t back_slot,
                  const mops_traj_io* io, mops_traj_stats* stats);

/* Asynchronous HOST-memory form (no counterpart in the reference, whose wrappers block: VK:1005-1014).  submit copies the
 * seeds up on a copy stream, enqueues the kernels and the copy-back (end points first, then the recorded trajectories) on a
 * second copy stream, and returns a ticket; two device staging sets alternate, so a second submit before the first wait
 * overlaps its H2D and the first call's D2H with the kernels (at most two tickets are in flight: a third submit first
 * completes the oldest).  The io buffers must stay valid (and should be pinned, mops_host_alloc) until the wait.
 * mops_traj_wait(what = 0) returns once io.xyz / io.depth hold the end points (what the next interval of a chain needs),
 * what = 1 once every output has landed; stats (may be NULL) are filled by the what = 1 wait. */
int mops_streamline_submit(mops_ctx* ctx, const mops_traj_cfg* cfg, int32_t slot, const mops_traj_io* io, int64_t* ticket);
int mops_pathline_submit(mops_ctx* ctx, const mops_traj_cfg* cfg, int32_t front_slot, int32_t back_slot,
                         const mops_traj_io* io, int64_t* ticket);
int mops_traj_wait(mops_ctx* ctx, int64_t ticket, int32_t what, mops_traj_stats* stats);

/* host-side line assembly + NaN trimming (src/Common/TrajectoryCommon.h:43-190; R7, R8):
 * raw [n][each][3] -> points / velocity [n][each+1][3], temperature / salinity [n][each+1]
 * (may be NULL), last [n][3].  pathline_mode fills temperature/salinity as the reference
 * does (from velocity.x / velocity.y).  Pure host function. */
int mops_finalize_lines(int64_t n, int32_t each, const double* seeds, const double* raw_pos, const double* raw_vel,
                        int32_t pathline_mode, double* points, double* velocity, double* temperature,
                        double* salinity, double* last);

/* ---- remap ---------------------------------------------------------------------------- */
typedef struct mops_remap_cfg {
    int32_t width, height;      /* VisualizationSettings::imageSize (x = width, y = height)   */
    double lat_min, lat_max;    /* LatRange */
    double lon_min, lon_max;    /* LonRange */
    double fixed_depth;         /* FixedDepth (positive down)                                 */
    int32_t mem;                /* MOPS_MEM_* for img0 / img1 / pixel_cell                    */
    int32_t reserved[3];
} mops_remap_cfg;

typedef struct mops_remap_stats {
    double kernel_ms;           /* locate + interpolate kernel                                */
    double total_ms;
    int64_t nan_pixels;
    int32_t launches;
    int32_t n_images;           /* 1, or 2 when the attribute image is produced               */
} mops_remap_stats;

/* replaces MOPS::Factory::VisualizeFixedDepth incl. its host KD-tree loop
 * (src/Common/MOPSFactory.h:15-19 -> VK:238-471, TBBKernel::SearchKDTree).
 * img0/img1: [height][width][4] doubles; img1 / pixel_cell may be NULL. */
int mops_remap_fixed_depth(mops_ctx* ctx, const mops_remap_cfg* cfg, int32_t slot,
                           double* img0, double* img1, int32_t* pixel_cell, mops_remap_stats* stats);

/* ---- the two other views of the reference ---------------------------------- */
typedef struct mops_view_cfg {
    int32_t width, height;      /* VisualizationSettings::imageSize                                   */
    double lat_min, lat_max;    /* fixed layer: LatRange                                              */
    double lon_min, lon_max;    /* LonRange (fixed latitude: inclusive end points, VK:513-514)        */
    int32_t fixed_layer;        /* fixed layer: FixedLayer, clamped to [0, nLevels-1] as ClampLayer   */
    int32_t mem;                /* MOPS_MEM_* of img / pixel_cell                                     */
    double fixed_latitude;      /* fixed latitude [deg]                                               */
    double depth_min, depth_max;/* fixed latitude: refBottomDepth.front() / .back() (row 0 / last row) */
} mops_view_cfg;
/* replaces MOPS::Factory::VisualizeFixedLayer (src/Common/MOPSFactory.h:9-13 -> VK:141-236) */
int mops_remap_fixed_layer(mops_ctx* ctx, const mops_view_cfg* cfg, int32_t slot, double* img, int32_t* pixel_cell,
                           mops_remap_stats* stats);
/* replaces MOPS::Factory::VisualizeFixedLatitude = MOPSApp::runReGrid
 * (src/Common/MOPSFactory.h:21-26 -> VK:473-651; host loops only in the reference's CUDA backend) */
int mops_regrid_fixed_latitude(mops_ctx* ctx, const mops_view_cfg* cfg, int32_t slot, double* img, int32_t* pixel_cell,
                               mops_remap_stats* stats);

/* ---- multi-GPU -------------------------------------------------------------------------------------
 * Particles shard, the mesh and the resident snapshots are replicated on every GPU, there is no collective inside the
 * time loop; the one exchange of the path is the gather of the recorded trajectories and end points to one owner, in
 * caller order (lineID = input index, src/Common/TrajectoryCommon.h:47,124), over NCCL (NVLink / NVSwitch).  The
 * reference has no counterpart: its only multi-rank code is a serial loop over MPI ranks in CLI/main.cpp:58-66,276-284.
 * NCCL is loaded at run time (libnccl.so.2); without it these calls return MOPS_E_STATE and the single-GPU ABI is unaffected. */

/* Processing-order key of a point: the rank of its cell along the mesh's Morton curve (what the engine sorts particles
 * by), -1 where the point has no cell.  Sharding = sort the seed set by this key and cut it into equal contiguous blocks
 * (mops_shard_bounds): equal counts for ANY seed distribution, spatially compact working set per GPU. */
int mops_order_key(mops_ctx* ctx, int32_t mem, int64_t n, const double* xyz, int32_t* key_out);
void mops_shard_bounds(int64_t n_total, int32_t rank, int32_t world, int64_t* lo, int64_t* hi);
void* mops_get_stream(mops_ctx* ctx); /* the cudaStream_t the context's kernels run on */
int mops_get_device(mops_ctx* ctx);

/* One process per GPU (torchrun, MPI ...): rank 0 obtains an id (mops_dist_unique_id, 128 bytes = ncclUniqueId), the caller
 * broadcasts it by its own means, every rank calls mops_dist_create with its context (collective). */
typedef struct mops_dist mops_dist;
#define MOPS_DIST_ID_BYTES 128
int mops_dist_unique_id(void* id128);
int mops_dist_create(mops_dist** out, mops_ctx* ctx, int32_t rank, int32_t world, const void* id128);
void mops_dist_destroy(mops_dist* d);
const char* mops_dist_last_error(const mops_dist* d);
/* run the gathers on a caller-owned cudaStream_t instead of the context's stream (NULL restores it), e.g. to overlap the
 * exchange of interval i with the kernels of interval i+1; ordering against the kernels is then the caller's (events) */
int mops_dist_set_stream(mops_dist* d, void* cuda_stream);
/* Collective.  Every rank contributes n_local = counts[rank] particles: index[i] = caller (global) index of local particle i,
 * pos / vel = [n_local][each][3] records, xyz = [n_local][3] end points, depth = [n_local] (any of the four may be NULL on
 * all ranks).  On `root` the rows land at their caller index in out_pos / out_vel [n_total][each][3], out_xyz [n_total][3],
 * out_depth [n_total].  All pointers are DEVICE pointers; variable-size ncclSend / ncclRecv in one group, enqueued on
 * the context's stream (asynchronous: order later work after it on that stream). */
int mops_dist_gather_traj(mops_dist* d, int32_t root, int64_t n_local, const int64_t* counts, const int32_t* index, int32_t each,
                          const double* pos, const double* vel, const double* xyz, const float* depth, int64_t n_total,
                          double* out_pos, double* out_vel, double* out_xyz, float* out_depth);

/* One process, N GPUs: N contexts + one host thread per device behind one handle (n_devices <= 0: every device of the box;
 * devices = NULL: ordinals 0..n-1).  Mesh and snapshots are replicated (each device pulls a snapshot over its own PCIe link,
 * in parallel); a trajectory call takes HOST buffers in caller order, sorts the seeds along the mesh's Morton curve on
 * device 0, cuts N equal contiguous blocks, ships each block to its device over NCCL, integrates the blocks concurrently
 * and gathers records / end points / status back to device 0 in caller order over NCCL (mops_dist_gather_traj's routine)
 * before the copy to the caller's buffers -- results are identical to the single-device call, whatever N.  This is what the
 * C++ drop-in's MOPS_RunStreamLine / MOPS_RunPathLine run on when more than one device is selected (MOPS_DEVICES=<n>).
 * Pixel views (remap, fixed layer / latitude) and point location run on mops_multi_ctx(m, 0). */
typedef struct mops_multi mops_multi;
int mops_multi_create(mops_multi** out, int32_t n_devices, const int32_t* devices);
void mops_multi_destroy(mops_multi* m);
const char* mops_multi_last_error(const mops_multi* m);
int32_t mops_multi_device_count(const mops_multi* m);
mops_ctx* mops_multi_ctx(mops_multi* m, int32_t i);
int mops_multi_set_mesh(mops_multi* m, int32_t n_cells, int32_t n_vertices, int32_t max_edges, const double* cell_xyz,
                        const double* vertex_xyz, const int32_t* vertices_on_cell, const int32_t* cells_on_cell,
                        const int32_t* cells_on_vertex, const int32_t* n_edges_on_cell);
int mops_multi_set_snapshot(mops_multi* m, int32_t slot, int32_t n_levels, const double* zonal, const double* meridional,
                            const double* layer_thickness, const double* bottom_depth, const double* vert_vel_top,
                            int32_t n_attr, const double* const* attrs, int32_t n_attr_total, int32_t async);
int mops_multi_snapshot_wait(mops_multi* m, int32_t slot);
int mops_multi_streamline(mops_multi* m, const mops_traj_cfg* cfg, int32_t slot, const mops_traj_io* io, mops_traj_stats* stats);
int mops_multi_pathline(mops_multi* m, const mops_traj_cfg* cfg, int32_t front_slot, int32_t back_slot, const mops_traj_io* io,
                        mops_traj_stats* stats);

/* ---- introspection ---------------------------------------------------------------- */
typedef struct mops_info {
    int32_t device, sm_count, cc_major, cc_minor;
    int64_t l2_bytes, hbm_bytes;
    int64_t mesh_bytes, snapshot_bytes[MOPS_MAX_SNAPSHOT_SLOTS];
    int32_t record_width;       /* compile-time vertex capacity of the resident cell records  */
    int32_t n_levels;
    int64_t total_launches;     /* kernels of this library launched since mops_create         */
    int32_t nonmonotone_cells[MOPS_MAX_SNAPSHOT_SLOTS]; /* cells taking the full-column path  */
} mops_info;
int mops_get_info(mops_ctx* ctx, mops_info* out);

#ifdef __cplusplus
}
#endif
#endif /* MOPS_B200_H */
