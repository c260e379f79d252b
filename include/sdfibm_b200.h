/*
 * sdfibm_b200.h — C ABI of the B200-native solid–fluid coupling path.
 *
 * This is the drop-in boundary for the ONE hot path of ChenguangZhang/sdfibm that this
 * repository re-implements on sm_100a: SolidCloud::interact (reference
 * src/solidcloud.cpp:435-464) with solidFluidInteract (:361-433), CellEnumerator
 * (src/cellenumerator.cpp:6-78), GeometricTools (src/geometrictools.cpp:6-116), the shape
 * SDFs (src/libshape/), checkAlpha (:564-570), fixInternal (:288-301) and the collision
 * step (:477-519, src/libcollision/).
 *
 * The reference has no FFI for this path: the seam is the C++ class sdfibm::SolidCloud
 * (src/solidcloud.h:89-123).  The host façade in sdfibm_b200/host/ keeps that class shape
 * and calls only the entry points below.  Plain pointers and sizes, no C++/torch types.
 *
 * Conventions
 *   - scalar = double, label = int32_t (reference CMakeLists.txt:20: -DWM_DP -DWM_LABEL_SIZE=32)
 *   - every entry returns 0 on success, non-zero on error; sdfibm_last_error() gives the
 *     thread-local message.  The library never calls exit().
 *   - there is NO CPU fallback: every compute entry fails if no CUDA device is usable.
 */
#ifndef SDFIBM_B200_H
#define SDFIBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFIBM_OK 0
#define SDFIBM_ERR_ARG 1
#define SDFIBM_ERR_CUDA 2
#define SDFIBM_ERR_STATE 3
#define SDFIBM_ERR_CAPACITY 4
#define SDFIBM_ERR_UNSUPPORTED 5

/* Shape tags of the device tagged union.  0..2 follow SHAPE2ID (src/libcollision/collision.h:11-15). */
enum sdfibm_shape_tag {
    SDFIBM_SHAPE_PLANE = 0,          /* src/libshape/plane.h:13-28 */
    SDFIBM_SHAPE_CIRCLE = 1,         /* src/libshape/circle.h:15-46 */
    SDFIBM_SHAPE_SPHERE = 2,         /* src/libshape/sphere.h:14-45 */
    SDFIBM_SHAPE_ELLIPSE = 3,        /* src/libshape/ellipse.h:14-54 */
    SDFIBM_SHAPE_ELLIPSOID = 4,      /* src/libshape/ellipsoid.h:15-54 */
    SDFIBM_SHAPE_RECTANGLE = 5,      /* src/libshape/rectangle.h:16-58 */
    SDFIBM_SHAPE_BOX = 6,            /* src/libshape/box.h:14-57 */
    SDFIBM_SHAPE_CIRCLE_TAIL = 7,    /* src/libshape/circle_tail.h:18-61 */
    SDFIBM_SHAPE_CIRCLE_TWOTAIL = 8, /* src/libshape/circle_twotail.h:18-66 */
    SDFIBM_SHAPE_PROGRAM = 9,        /* a composed shape: a post-fix program over src/libshape/sdf/sdf.h (see sdfibm_sdf_op_t) */
    SDFIBM_SHAPE_NTAGS = 10
};

/* Composable signed-distance programs: the primitives, point transformations and Boolean operations of the reference's
 * sdf:: namespace (src/libshape/sdf/sdf.h:13-155) — what a plugin written from src/libshape/template.h combines in its
 * isInside / signedDistance pair — as a post-fix program the device evaluates, so a NEW composed shape needs no change to the
 * CUDA switch.  Two stacks: points and values (a value = the bool of isInside and the scalar of signedDistance, computed
 * side by side with the reference's own expressions: `<` predicates un-contracted).
 *   POINT / POINT_2D          push  com + p   (POINT_2D: z = 0, the 2-D shapes' p2d)
 *   OFFSET a0 a1 a2           top point -= (a0, a1, a2)                                      sdf::offset      :123-126
 *   ROT30/45/60/90, ROTTH a0  rotate the top point (the literals of sdf.h, not exact cosines)  sdf::rot*       :91-113
 *   FLIPX / FLIPY             mirror the top point                                           sdf::flipx/y     :115-122
 *   CIRCLE a0=r a1=r^2        pop point, push (|P|^2 < a1, |P| - a0)   (a sphere when the point is 3-D)       :15-26
 *   RECTANGLE a0=ra a1=rb     sdf::rectangle_bool / rectangle                                                :29-40
 *   BOX a0 a1 a2              sdf::box_bool / box                                                            :43-56
 *   ELLIPSE a0=1/a^2 a1=1/b^2, ELLIPSOID a0 a1 a2 = 1/a^2 1/b^2 1/c^2                                         :59-83
 *   HALFSPACE                 (P.y < 0, P.y)                                                 plane.h:21-28
 *   UNION / INTERSECT / DIFF  pop two values, push sdf::U / I / D of them (n-ary U({..}) = repeated UNION)   :131-142
 * The program must leave exactly one value; the shape's signedDistance is sdf::filter of it (:147-150). */
enum sdfibm_sdf_opcode {
    SDFIBM_OP_POINT = 0, SDFIBM_OP_POINT_2D = 1, SDFIBM_OP_OFFSET = 2,
    SDFIBM_OP_ROT30 = 3, SDFIBM_OP_ROT45 = 4, SDFIBM_OP_ROT60 = 5, SDFIBM_OP_ROT90 = 6, SDFIBM_OP_ROTTH = 7,
    SDFIBM_OP_FLIPX = 8, SDFIBM_OP_FLIPY = 9,
    SDFIBM_OP_CIRCLE = 16, SDFIBM_OP_RECTANGLE = 17, SDFIBM_OP_BOX = 18, SDFIBM_OP_ELLIPSE = 19, SDFIBM_OP_ELLIPSOID = 20,
    SDFIBM_OP_HALFSPACE = 21,
    SDFIBM_OP_UNION = 32, SDFIBM_OP_INTERSECT = 33, SDFIBM_OP_DIFF = 34
};
#define SDFIBM_SDF_STACK 6   /* depth of either stack */
typedef struct sdfibm_sdf_op {
    int32_t op;
    int32_t pad_;
    double a[4];
} sdfibm_sdf_op_t;

/* Cell classification, same numbering as CellEnumerator::CELL_TYPE (src/cellenumerator.h:23). */
enum sdfibm_cell_type {
    SDFIBM_CELL_UNVISITED = 0,
    SDFIBM_CELL_ALL_INSIDE = 1,
    SDFIBM_CELL_CENTER_INSIDE = 2,
    SDFIBM_CELL_CENTER_OUTSIDE = 3,
    SDFIBM_CELL_ALL_OUTSIDE = 4
};

/*
 * One registered shape, lowered to a POD record (what IShape subclasses keep as members).
 * p[] holds the members each reference class derives in its constructor:
 *   PLANE           -
 *   CIRCLE, SPHERE  p0 = radius, p1 = radius*radius
 *   ELLIPSE         p0 = a, p1 = b, p2 = 1/(a*a), p3 = 1/(b*b)
 *   ELLIPSOID       p0..2 = a,b,c, p3..5 = 1/(a*a), 1/(b*b), 1/(c*c)      (ignores com)
 *   RECTANGLE       p0 = a, p1 = b (half widths)
 *   BOX             p0..2 = a,b,c
 *   CIRCLE_TAIL     p0 = radius, p1 = radius^2, p2 = (ratio+1)*0.5*radius, p3 = thickness
 *   CIRCLE_TWOTAIL  p0 = radius, p1 = radius^2, p2 = (ratio+1)*0.5*radius, p3 = thickness*0.5
 */
typedef struct sdfibm_shape {
    int32_t tag;      /* enum sdfibm_shape_tag */
    int32_t finite;   /* IShape::finite (src/libshape/ishape.h:38) */
    double radiusB;   /* IShape::m_radiusB — used by the narrow phase (collision.cpp:12-13) */
    double com[3];    /* IShape::m_com, ADDED to the local point (circle.h:40) */
    double p[8];
} sdfibm_shape_t;

/* Rigid-body state the path reads (src/solid.h:17-28).  quat = (w, x, y, z). */
typedef struct sdfibm_solid {
    double pos[3];
    double quat[4];
    double vel[3];
    double omega[3];
    int32_t shape; /* index into the shape table */
    int32_t pad_;
} sdfibm_solid_t;

/*
 * The mesh arrays MeshInfo binds (src/meshinfo.h:20-29) plus mesh.cells()/mesh.faces()
 * (src/geometrictools.cpp:61,66), in CSR form, local (per-rank) numbering.
 */
typedef struct sdfibm_mesh {
    int32_t n_cells, n_points, n_faces, n_internal_faces;
    const double *points;        /* [n_points*3]  mesh.points()      */
    const double *cell_centres;  /* [n_cells*3]   mesh.cellCentres() */
    const double *cell_volumes;  /* [n_cells]     mesh.V()           */
    const double *face_centres;  /* [n_faces*3]   mesh.faceCentres() */
    const double *face_areas;    /* [n_faces*3]   mesh.faceAreas()   */
    const int32_t *cell_points_off; /* [n_cells+1] */
    const int32_t *cell_points;     /* cellPoints(): ascending point label per cell */
    const int32_t *cell_faces_off;  /* [n_cells+1] */
    const int32_t *cell_faces;      /* cells(): owned faces ascending, then neighbour faces */
    const int32_t *face_points_off; /* [n_faces+1] */
    const int32_t *face_points;     /* faces(): vertex loop as stored */
    const int32_t *cell_cells_off;  /* [n_cells+1] */
    const int32_t *cell_cells;      /* cellCells(): ascending internal-face order */
    double bounds_min[3], bounds_max[3]; /* mesh.bounds() */
} sdfibm_mesh_t;

typedef struct sdfibm_context sdfibm_context;

/* ---- life cycle -------------------------------------------------------------------- */
int sdfibm_version(void);
const char *sdfibm_last_error(void);
int sdfibm_device_count(int *count);
/* One context per rank/GPU; owns one CUDA stream.  replaces SolidCloud ctor device part (solidcloud.cpp:209-217). */
int sdfibm_create(int device, sdfibm_context **ctx);
int sdfibm_destroy(sdfibm_context *ctx);
/* Page-locked host memory for the per-step arrays the host side owns (solid states, U, the four fields): copies from / to
 * such memory are asynchronous and skip the library's staging copy.  Optional — any host pointer is accepted everywhere. */
int sdfibm_alloc_pinned(size_t bytes, void **out);
int sdfibm_free_pinned(void *p);
/* max solids that may touch one mesh cell (default 3); call before sdfibm_set_mesh */
int sdfibm_set_cell_slots(sdfibm_context *ctx, int slots);

/* Meshes whose cells differ in vertex count (hanging-node refinement, hex / prism / polyhedron mixes): the reference's
 * ALL_INSIDE test compares a cell's inside-vertex count with the vertex count of the cell that discovered it in the flood fill
 * (src/cellenumerator.cpp:25), i.e. it depends on the fill's visiting order.  That order is not reproduced: sdfibm_set_mesh
 * refuses such a mesh (SDFIBM_ERR_UNSUPPORTED) unless the caller accepts the order-free rule — ALL_INSIDE when all of the cell's
 * OWN vertices are inside — here (or with SDFIBM_ALLOW_ORDER_FREE=1).  Member cells, fractions of cut cells and forces are
 * unaffected; only the type (and hence alpha = 1 vs the computed fraction) of a few cells next to a cell of another kind is. */
int sdfibm_allow_order_free(sdfibm_context *ctx, int on);

/* ---- one-time uploads ------------------------------------------------------------- */
/* replaces GeometricTools(mesh)/MeshInfo(mesh) binding (solidcloud.cpp:216, meshinfo.h:20-29) */
int sdfibm_set_mesh(sdfibm_context *ctx, const sdfibm_mesh_t *mesh, int two_d);
/* replaces EntityLibrary<IShape> lookup by pointer (solidcloud.cpp:59, solid.h:47-50) */
int sdfibm_set_shapes(sdfibm_context *ctx, const sdfibm_shape_t *shapes, int n_shapes);
/* The op table the SDFIBM_SHAPE_PROGRAM records of the shape table point into; call BEFORE sdfibm_set_shapes.  A PROGRAM record:
 * p[0] = index of its first op, p[1] = number of ops, p[2] = certified outer radius about the body origin (no point farther is
 * inside; in-plane for 2-D), p[3] = certified inner radius (every point closer is inside; 0 if unknown), p[4] = 1 for a 2-D shape
 * (unbounded along the body z axis), else 0.  Programs are checked (stack discipline, opcodes) when the shapes are set. */
int sdfibm_set_shape_programs(sdfibm_context *ctx, const sdfibm_sdf_op_t *ops, int n_ops);

/* ---- SolidCloud::interact (solidcloud.cpp:435-464), host buffers --------------------
 * U[3*n_cells] in;  As[n_cells], Fs[3*n_cells], Ts[n_cells], Ct[n_cells] out (cell values of
 * the registered volFields);  force_torque[6*n_solids] out = per-solid (F, T) already
 * multiplied by rhof (solidcloud.cpp:424-425), BEFORE any cross-rank reduction. */
int sdfibm_interact(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids,
                    const double *U, double dt, double rhof,
                    double *As, double *Fs, double *Ts, double *Ct, double *force_torque);
/* Same with DEVICE pointers on the context's device, enqueued on the context stream and
 * synchronised before return (the per-step status word needs one 4-byte read-back). */
int sdfibm_interact_device(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids,
                           const double *dU, double dt, double rhof,
                           double *dAs, double *dFs, double *dTs, double *dCt,
                           double *d_force_torque);

/* Same, with the solid records ALREADY on the context's device (enqueued before this call on the context stream, e.g. one
 * 1/N slice uploaded per rank and all-gathered over NVLink instead of N identical PCIe uploads of the replicated state).
 * may_be_global = 0 promises that no solid is a plane or a 2-D shape whose axis is not exactly world z (only consulted when the
 * shape table holds such shapes). */
int sdfibm_interact_device_solids(sdfibm_context *ctx, const sdfibm_solid_t *d_solids, int n_solids, int may_be_global,
                                  const double *dU, double dt, double rhof,
                                  double *dAs, double *dFs, double *dTs, double *dCt,
                                  double *d_force_torque);

/* ---- SolidCloud::fixInternal (solidcloud.cpp:288-301) -------------------------------
 * Uses Ct of the last interact on this context and the CURRENT solid states. */
int sdfibm_fix_internal(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, double *U);
int sdfibm_fix_internal_device(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids,
                               double *dU, const double *dCt);

/* ---- the step either side of interact, on the device (src/main.cpp:70-77) ---------------
 * With the flow fields resident in HBM (sdfibm_interact_device), the forcing is applied where it was computed:
 *   U = U - Fs dt   (main.cpp:70)      T = (1 - As) T + Ts   (main.cpp:76)
 * using As / Fs / Ts of the LAST interact on this context (their device arrays must still be alive).  Only cells that hold a
 * candidate record are touched — everywhere else both updates are the identity bit for bit — so the pass moves 100-odd bytes
 * per touched cell instead of 120 bytes per mesh cell.  Either pointer may be NULL.  Stream-ordered (no host synchronisation).
 * Per step the host then sends the solid states (112 B each) and reads back force_torque (48 B per solid): the 1.2 GB of
 * PCIe traffic of the host-buffer entry at C4 becomes 1.6 MB. */
int sdfibm_apply_forcing_device(sdfibm_context *ctx, double *dU, double *dT, double dt);
/* Copy `bytes` from a device array to host memory behind everything enqueued on the context stream, and wait for it (the one
 * host synchronisation of a device-resident step: e.g. force_torque after sdfibm_interact_device + sdfibm_apply_forcing_device). */
int sdfibm_download(sdfibm_context *ctx, void *host_dst, const void *device_src, size_t bytes);
/* Compact records of the cells the last interact touched (a solid's bounding volume reached them; every other cell of the four
 * fields is zero): cells[n] = cell labels, As[n], Fs[3n], Ts[n], Ct[n] (host arrays).  cells == NULL: only *n_touched is
 * returned (size the arrays with it).  For a host that keeps its own copy of the fields and wants the D2H traffic of a step to
 * scale with the solids, not with the mesh. */
int sdfibm_touched_cells(sdfibm_context *ctx, int64_t capacity, int64_t *n_touched,
                         int32_t *cells, double *As, double *Fs, double *Ts, double *Ct);

/* ---- mean-field sampler: SolidCloud::calcMeanField (solidcloud.cpp:315-359) ------------
 * For every solid (placed with the SUBSTITUTE shape its record names) the volume-weighted mean of a cell field over
 * the shape's candidate cells: mean[3*s..] = sum(alpha V field) / sum(alpha V); sum_alpha_v[s] = the denominator
 * (before any cross-rank reduction; pass NULL if not wanted).  field[3*n_cells] is a host array.  Runs the interact
 * kernels on scratch outputs: the fields and Ct of the last sdfibm_interact stay valid for sdfibm_fix_internal
 * (the reference's sampler overwrites Ct as a side effect, SURVEY Q11 — not reproduced). */
int sdfibm_mean_field(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, const double *field,
                      double *mean, double *sum_alpha_v);
/* The raw sums of the same sampler: sum_alpha_v_field[3*s..] = sum(alpha V field), sum_alpha_v[s] = sum(alpha V) over THIS
 * rank's cells.  A parallel host reduces both across ranks and divides afterwards (solidcloud.cpp:353-357) — a rank whose
 * block does not touch the solid contributes zeros, not 0/0. */
int sdfibm_mean_field_sums(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, const double *field,
                           double *sum_alpha_v_field, double *sum_alpha_v);

/* ---- tool_vof: SolidCloud::writeVOF (tool_vof/solidcloud.cpp:116-173) ----------------------------------------
 * alpha[n_cells] (host) = the volume fraction every cell has inside the union of the solids: the per-solid fractions added
 * in solid order and clamped at 1 (:131-133) — interact's As, computed by the same kernels on scratch outputs (the coupling
 * state of the last interact on this context stays valid).  The reference lists its `planes{}` block after `solids{}`; pass
 * them in that order.  total_volume (may be NULL) = sum(alpha V) (:169). */
int sdfibm_volume_fraction(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, double *alpha, double *total_volume);

/* ---- candidate lists of the last interact (CellEnumerator::intersect result) --------
 * counts[3] = total ALL_INSIDE, CENTER_INSIDE, CENTER_OUTSIDE pairs. */
int sdfibm_candidate_counts(sdfibm_context *ctx, int64_t counts[3]);
/* offsets[3*n_solids+1]: segment (3*s + type-1) of cells[] holds the ascending cell ids of
 * solid s and type (1,2,3) — std::set order (cellenumerator.h:25, solidcloud.cpp:367-374). */
int sdfibm_candidate_lists(sdfibm_context *ctx, int32_t *offsets, int32_t *cells, int64_t capacity);
/* diagnostics of the last interact: [0] solids whose vertex-inside cell set was not one
 * face-connected component (exact flood-fill replay was run), [1] launches enqueued, [2] solid-bin
 * entries, [3] (cell, solid) items that needed exact vertex evaluation. */
int sdfibm_last_stats(sdfibm_context *ctx, int64_t stats[4]);
/* device time (CUDA events on the context stream) of the last interact, in ms:
 * [0] solid preparation + binning, [1] k_classify, [2] k_heavy, [3] k_final, [4] connectivity + finalise,
 * [5] whole pipeline */
int sdfibm_last_timings(sdfibm_context *ctx, double ms[6]);
/* host wall time of the last interact, in microseconds: [0] staging the solid states, [1] enqueue / graph launch,
 * [2] waiting for the GPU, [3] the whole call */
int sdfibm_last_host_timings(sdfibm_context *ctx, double us[4]);
/* device time (CUDA events on the context stream, ms) of the kernels of [0] the last sdfibm_fix_internal_device and [1] the last
 * sdfibm_collide (key build, sort, pair enumeration, narrow phase; the pair-count read-back in the middle included) */
int sdfibm_last_aux_timings(sdfibm_context *ctx, double ms[2]);

/* ---- collision step (solidcloud.cpp:477-519, libcollision/) -------------------------
 * delta = UGrid cell size.  HEAD passes 2*m_radiusB = -2 (solidcloud.cpp:74-75,245) which yields
 * no pairs; pass a positive delta for the intended behaviour.  pairs[2*cap] receives (p<q)
 * pairs in UGrid::generateCollisionPairs order; force_torque[6*n_solids] is ACCUMULATED into. */
int sdfibm_collide(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, double delta,
                   int32_t *pairs, int64_t pair_capacity, int64_t *n_pairs, double *force_torque);

/* ---- device-resident access for the multi-GPU host --------------------------------- */
int sdfibm_stream(sdfibm_context *ctx, void **cuda_stream);
int sdfibm_synchronize(sdfibm_context *ctx);

/* ---- cross-rank exchange: one process per GPU, cells partitioned (decomposePar), solids replicated ----
 * Replaces the 2N Foam::reduce calls of src/solidcloud.cpp:427-431 by ONE ncclAllReduce(sum) of 6N fp64 on the context
 * stream.  NCCL is bound at run time (dlopen of libnccl.so.2): a serial host never needs it.
 *   rank 0:      sdfibm_comm_unique_id(id)  ->  the host broadcasts the 128 bytes (MPI_Bcast / Pstream::scatter / a file)
 *   every rank:  sdfibm_comm_init(ctx, id, rank, n_ranks)      (collective)
 * After sdfibm_comm_init every sdfibm_interact / sdfibm_interact_device on the context
 *   - uploads only this rank's 1/N slice of the (replicated) solid array and all-gathers the slices over NVLink, and
 *   - returns force_torque already summed over the ranks (on the device-resident entries the all-reduce runs on a communication
 *     stream of the library alongside the last kernels of the step; a rank that has to re-run its step — queue growth,
 *     flood-fill replay — tells the others through a flag that is reduced with it, and every rank then reduces again).
 * Both calls are COLLECTIVE from then on: every rank must make them with the same solid array.  sdfibm_comm_options switches
 * either behaviour off (e.g. to reduce by MPI instead); sdfibm_allreduce_force_torque is the bare collective on a device
 * array, in place, stream-ordered (no host synchronisation). */
#define SDFIBM_COMM_ID_BYTES 128
int sdfibm_comm_unique_id(void *id /* [SDFIBM_COMM_ID_BYTES] */);
int sdfibm_comm_init(sdfibm_context *ctx, const void *id, int rank, int n_ranks);
int sdfibm_comm_options(sdfibm_context *ctx, int auto_reduce, int gather_solids);
int sdfibm_comm_destroy(sdfibm_context *ctx);
int sdfibm_allreduce_force_torque(sdfibm_context *ctx, double *d_force_torque, int n_solids);
/* device time (ms) of the collectives of the last interact on this context */
int sdfibm_comm_last_ms(sdfibm_context *ctx, double *ms);

/* ---- Foam-free mesh helpers (stand-alone harness; OpenFOAM supplies these in the drop-in) */
/* derived geometry + connectivity from polyMesh primitives (points/faces/owner/neighbour) */
typedef struct sdfibm_mesh_storage sdfibm_mesh_storage;
int sdfibm_mesh_from_polymesh(int32_t n_points, const double *points, int32_t n_faces,
                              const int32_t *face_off, const int32_t *face_pts,
                              const int32_t *owner, int32_t n_internal, const int32_t *neighbour,
                              sdfibm_mesh_storage **out);
/* single blockMesh-numbered hex block: nx*ny*nz cells, origin x0, spacing dx */
int sdfibm_mesh_hex_block(int32_t nx, int32_t ny, int32_t nz, const double x0[3], const double dx[3],
                          sdfibm_mesh_storage **out);
int sdfibm_mesh_view(const sdfibm_mesh_storage *st, sdfibm_mesh_t *view);
/* polyMesh owner[n_faces] / neighbour[n_internal_faces] of the stored mesh */
int sdfibm_mesh_owner_neighbour(const sdfibm_mesh_storage *st, const int32_t **owner, const int32_t **neighbour);
int sdfibm_mesh_free(sdfibm_mesh_storage *st);

#ifdef __cplusplus
}
#endif
#endif /* SDFIBM_B200_H */
