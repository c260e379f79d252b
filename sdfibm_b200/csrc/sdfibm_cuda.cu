// sdfibm_cuda.cu — sm_100a kernels and the C-ABI implementation of the coupling path.
//
// Design: the reference walks solids and flood-fills cells per solid
// (src/solidcloud.cpp:445-446, src/cellenumerator.cpp:6-34).  Here the loop is inverted: ONE thread
// per mesh cell walks the few solids whose (conservatively inflated) bounding volume covers that
// cell, taken from a per-step uniform-grid binning of the solids.  Every field (As, Fs, Ts, Ct) is
// then written exactly once per cell, fully coalesced, with the per-cell sums taken in ascending
// solid order exactly like the reference's `+=` over its solid loop — no field memsets, no field
// atomics.  The flood fill's "face-connected component of the seed" semantics is
// restored by a connectivity certificate kernel and, only for solids that fail it, an exact
// label-propagation replay (all on the GPU).
//
// Compiled with -fmad=false: predicates are strict `<` on un-contracted fp64.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "device_math.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void sdfibm_set_error(const std::string &msg) { g_last_error = msg; }
static int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(SDFIBM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device-side records
// ------------------------------------------------------------------------------------------------
#define KIND_3D 0      // bounded in all directions about pos
#define KIND_2D 1      // infinite along the body z axis (z is zeroed by the shape)
#define KIND_PLANE 2   // half space
#define MAX_CELL_VERTS 32
#define MAX_FACE_VERTS 16
#define REL_MARGIN 1e-6

struct DevShape {
    sdfibm_shape_t s;
    double r_out; // certified: no point farther than r_out (in-plane for 2-D) from the body origin is inside
    double r_in;  // certified: every point closer than r_in (3-D distance) to the body origin is inside
    int kind;
    int pad;
};

struct DevSolid {
    double pos[3];
    double q[4];
    double vel[3];
    double omega[3];
    double axis[3]; // world direction of the body z axis
    double r_out, r_in;
    int shape;
    int kind;
    int axis_is_z; // body z axis coincides with world z (2-D cases)
    int global;    // not binned: tested by every cell
};

struct DevMesh {
    int n_cells, n_points, n_faces;
    const double *points, *cc, *V, *Cf, *Sf;
    const int *cp_off, *cp, *cf_off, *cf, *fp_off, *fp, *nb_off, *nb;
    const float2 *cell_rad; // (3-D radius, xy radius) of the vertex cloud about the centre, rounded up
    const double *magSf;    // |Sf| per face (same expression as Foam::mag, evaluated once at upload)
    const unsigned *hex_topo; // hex meshes: 3 words per cell, 4-bit cell-local vertex slot of every face vertex
    float2 rad_const;       // upper bound of cell_rad over the mesh (used for every cell when the mesh is near uniform)
    int rad_uniform;
    int is_hex;             // every cell has 8 points / 6 faces and every face 4 points
    int two_d;
};

struct BinGrid {
    double lo[3];
    double inv_b;
    int n[3];
    int n_bins;
};

struct StepStatus {
    unsigned long long counts[3]; // ALL_INSIDE, CENTER_INSIDE, CENTER_OUTSIDE pairs
    int n_flagged;                // solids with >1 certificate roots
    int slot_overflow;            // a cell was touched by more than K solids
    int bin_overflow;             // bin list capacity exceeded
    int bad_cell;                 // a cell/face exceeded MAX_CELL_VERTS / MAX_FACE_VERTS
    int bin_total;
    int n_global;
    unsigned long long heavy_total; // (cell, solid) items that needed exact evaluation
};

__device__ __forceinline__ D3 ld3(const double *__restrict__ p, long long i) {
    return {__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)};
}

__device__ __forceinline__ int bin_coord(const BinGrid &g, double x, int d) {
    double t = floor((x - g.lo[d]) * g.inv_b);
    int i = (t < 0.0) ? 0 : ((t > (double)(g.n[d] - 1)) ? g.n[d] - 1 : (int)t);
    return i;
}

// ------------------------------------------------------------------------------------------------
// K0  per-cell vertex-cloud radii (once per mesh)
// ------------------------------------------------------------------------------------------------
__global__ void k_face_mag(const double *Sf, int n_faces, double *magSf) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n_faces) magSf[f] = mag3(ld3(Sf, f));
}

// hex meshes: for face k (cells() order) of cell c and its j-th vertex (faces() order), the position of that
// vertex in the cell's cellPoints() list, packed 4 bits each: word k/2, bit 16*(k&1) + 4*j.
__global__ void k_hex_topo(DevMesh m, unsigned *topo, int *bad) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.n_cells) return;
    int vid[8];
    for (int k = 0; k < 8; ++k) vid[k] = m.cp[8 * (long long)c + k];
    unsigned w[3] = {0u, 0u, 0u};
    for (int k = 0; k < 6; ++k) {
        const int f = m.cf[6 * (long long)c + k];
        for (int j = 0; j < 4; ++j) {
            const int g = m.fp[4 * (long long)f + j];
            int l = -1;
            for (int t = 0; t < 8; ++t) if (vid[t] == g) l = t;
            if (l < 0) { atomicExch(bad, 1); l = 0; }
            w[k >> 1] |= (unsigned)l << (16 * (k & 1) + 4 * j);
        }
    }
    topo[3 * (long long)c] = w[0];
    topo[3 * (long long)c + 1] = w[1];
    topo[3 * (long long)c + 2] = w[2];
}

// rmax[0..1] = max of the (3-D, xy) radii, rmax[2..3] = min
__global__ void k_cell_radius(DevMesh m, float2 *rad, int *bad, float *rmax) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    float r3f = 0.f, rxyf = 0.f;
    if (c < m.n_cells) {
        D3 cc = ld3(m.cc, c);
        double r3 = 0.0, rxy = 0.0;
        int b = m.cp_off[c], e = m.cp_off[c + 1];
        if (e - b > MAX_CELL_VERTS) atomicExch(bad, 1);
        for (int k = b; k < e; ++k) {
            D3 d = ld3(m.points, m.cp[k]) - cc;
            r3 = fmax(r3, d.x * d.x + d.y * d.y + d.z * d.z);
            rxy = fmax(rxy, d.x * d.x + d.y * d.y);
        }
        for (int k = m.cf_off[c]; k < m.cf_off[c + 1]; ++k) {
            int f = m.cf[k];
            if (m.fp_off[f + 1] - m.fp_off[f] > MAX_FACE_VERTS) atomicExch(bad, 1);
        }
        r3f = __double2float_ru(sqrt(r3) * (1.0 + REL_MARGIN));
        rxyf = __double2float_ru(sqrt(rxy) * (1.0 + REL_MARGIN));
        r3f = nextafterf(r3f, INFINITY);
        rxyf = nextafterf(rxyf, INFINITY);
        rad[c] = make_float2(r3f, rxyf);
    }
    // block max/min -> global (floats are non-negative: int compare is order preserving)
    __shared__ float s3[256], sxy[256], m3[256], mxy[256];
    s3[threadIdx.x] = r3f;
    sxy[threadIdx.x] = rxyf;
    m3[threadIdx.x] = (c < m.n_cells) ? r3f : 3.0e38f;
    mxy[threadIdx.x] = (c < m.n_cells) ? rxyf : 3.0e38f;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s3[threadIdx.x] = fmaxf(s3[threadIdx.x], s3[threadIdx.x + o]);
            sxy[threadIdx.x] = fmaxf(sxy[threadIdx.x], sxy[threadIdx.x + o]);
            m3[threadIdx.x] = fminf(m3[threadIdx.x], m3[threadIdx.x + o]);
            mxy[threadIdx.x] = fminf(mxy[threadIdx.x], mxy[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicMax((int *)&rmax[0], __float_as_int(s3[0]));
        atomicMax((int *)&rmax[1], __float_as_int(sxy[0]));
        atomicMin((int *)&rmax[2], __float_as_int(m3[0]));
        atomicMin((int *)&rmax[3], __float_as_int(mxy[0]));
    }
}

// ------------------------------------------------------------------------------------------------
// K1  per-step solid preparation + bin counting / filling
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool solid_bin_range(const DevSolid &S, const BinGrid &g, double rad3_max, double radxy_max,
                                                const double *mesh_lo, const double *mesh_hi, int lo[3], int hi[3]) {
    double ext[3];
    if (S.kind == KIND_3D) {
        ext[0] = ext[1] = ext[2] = S.r_out + rad3_max;
    } else { // KIND_2D with axis == world z: bounded in x,y only
        ext[0] = ext[1] = S.r_out + radxy_max;
        ext[2] = 1e300;
    }
    for (int d = 0; d < 3; ++d) {
        double a = S.pos[d] - ext[d], b = S.pos[d] + ext[d];
        if (b < mesh_lo[d] || a > mesh_hi[d]) return false;
        lo[d] = bin_coord(g, a, d);
        hi[d] = bin_coord(g, b, d);
    }
    return true;
}

struct PrepParams {
    const sdfibm_solid_t *solids;
    const DevShape *shapes;
    int n_solids, n_shapes;
    DevSolid *out;
    BinGrid grid;
    double rad3_max, radxy_max;
    double mesh_lo[3], mesh_hi[3];
    int *bin_count;   // [n_bins+1]
    int *global_list; // [n_solids]
    StepStatus *status;
};

__global__ void k_solid_prepare(PrepParams P) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_solids) return;
    const sdfibm_solid_t in = P.solids[s];
    DevSolid S;
    for (int d = 0; d < 3; ++d) { S.pos[d] = in.pos[d]; S.vel[d] = in.vel[d]; S.omega[d] = in.omega[d]; }
    for (int d = 0; d < 4; ++d) S.q[d] = in.quat[d];
    S.shape = in.shape;
    const DevShape &sh = P.shapes[in.shape];
    S.kind = sh.kind;
    S.r_out = sh.r_out;
    S.r_in = sh.r_in;
    DQ q = {in.quat[0], {in.quat[1], in.quat[2], in.quat[3]}};
    D3 ax = qtransform(q, D3{0.0, 0.0, 1.0});
    S.axis[0] = ax.x; S.axis[1] = ax.y; S.axis[2] = ax.z;
    S.axis_is_z = (fabs(ax.x) <= 1e-12 && fabs(ax.y) <= 1e-12) ? 1 : 0;
    S.global = (S.kind == KIND_PLANE || (S.kind == KIND_2D && !S.axis_is_z)) ? 1 : 0;
    P.out[s] = S;
    if (S.global) {
        int k = atomicAdd(&P.status->n_global, 1);
        P.global_list[k] = s;
        return;
    }
    int lo[3], hi[3];
    if (!solid_bin_range(S, P.grid, P.rad3_max, P.radxy_max, P.mesh_lo, P.mesh_hi, lo, hi)) return;
    for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) atomicAdd(&P.bin_count[(k * P.grid.n[1] + j) * P.grid.n[0] + i], 1);
}

struct FillParams {
    const DevSolid *solids;
    int n_solids;
    BinGrid grid;
    double rad3_max, radxy_max;
    double mesh_lo[3], mesh_hi[3];
    const int *bin_off; // exclusive scan of bin_count, [n_bins+1]
    int *bin_cursor;    // zeroed, [n_bins]
    int *bin_list;
    int bin_cap;
    StepStatus *status;
};

__global__ void k_bin_fill(FillParams P) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_solids) return;
    const DevSolid S = P.solids[s];
    if (S.global) return;
    int lo[3], hi[3];
    if (!solid_bin_range(S, P.grid, P.rad3_max, P.radxy_max, P.mesh_lo, P.mesh_hi, lo, hi)) return;
    for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) {
                int b = (k * P.grid.n[1] + j) * P.grid.n[0] + i;
                int pos = P.bin_off[b] + atomicAdd(&P.bin_cursor[b], 1);
                if (pos < P.bin_cap) P.bin_list[pos] = s;
                else P.status->bin_overflow = 1;
            }
}

// ascending solid id inside every bin (and the global list): the per-cell accumulation order
__global__ void k_bin_sort(const int *bin_off, int *bin_list, int n_bins, int bin_cap, int *global_list, StepStatus *status) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    int beg, end;
    int *lst;
    if (b < n_bins) { beg = bin_off[b]; end = bin_off[b + 1]; lst = bin_list; if (end > bin_cap) return; }
    else if (b == n_bins) { beg = 0; end = status->n_global; lst = global_list; status->bin_total = bin_off[n_bins]; }
    else return;
    for (int i = beg + 1; i < end; ++i) {
        int v = lst[i], j = i - 1;
        while (j >= beg && lst[j] > v) { lst[j + 1] = lst[j]; --j; }
        lst[j + 1] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// geometry: apex / pyramid volume fraction (reference src/geometrictools.cpp:13-116)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double line_fraction(double a, double b) {          // :13-23
    if (a > 0 && b > 0) return 0.0;
    if (a <= 0 && b <= 0) return 1.0;
    if (a > 0) return -b / (a - b);
    return -a / (b - a);
}

// calcApex over an indexed vertex list (:25-45).  pts/phi are the cell-local arrays, idx maps the
// list position to the cell-local slot (identity for the cell's own vertex list).
template <typename IDX>
__device__ __forceinline__ D3 calc_apex(const D3 *pts, const double *phi, IDX idx, int n) {
    const int i0 = idx(0);
    D3 A = pts[i0];
    double phiA = phi[i0];
    D3 B = {0.0, 0.0, 0.0};
    double phiB = 0.0;
    for (int i = 1; i < n; ++i) {
        int ii = idx(i);
        B = pts[ii];
        phiB = phi[ii];
        if (phiA * phiB <= 0) break;
    }
    return A - fabs(phiA) / (SDF_SMALL + fabs(phiA) + fabs(phiB)) * (A - B);
}

// calcCellVolume (:47-72) for cell c with cell-local vertex coordinates and phi already evaluated.
__device__ double cell_solid_volume(const DevMesh &m, int c, const int *vid, const D3 *pts, const double *phi, int nv) {
    D3 apex = calc_apex(pts, phi, [](int i) { return i; }, nv);
    if (m.two_d) apex.z = 0.0;
    double volume = 0.0;
    const int fb = __ldg(m.cf_off + c), fe = __ldg(m.cf_off + c + 1);
    for (int k = fb; k < fe; ++k) {
        const int f = __ldg(m.cf + k);
        const int pb = __ldg(m.fp_off + f), nf = __ldg(m.fp_off + f + 1) - pb;
        int loc[MAX_FACE_VERTS];
        int sign_sum = 0;
        for (int i = 0; i < nf; ++i) {
            const int g = __ldg(m.fp + pb + i);
            int l = 0;
            while (l < nv - 1 && vid[l] != g) ++l;
            loc[i] = l;
            if (phi[l] > 0) ++sign_sum;
            else --sign_sum;
        }
        double eps_f;
        const D3 Sf = ld3(m.Sf, f);
        if (sign_sum == nf) eps_f = 0.0;                                        // :98-116
        else if (sign_sum == -nf) eps_f = 1.0;
        else {
            D3 fap = calc_apex(pts, phi, [&](int i) { return loc[i]; }, nf);    // calcFaceArea :74-96
            double area = 0.0;
            for (int i = 0; i < nf; ++i) {
                const int lo = loc[i], la = loc[(i + 1) % nf];
                const D3 O = pts[lo], A = pts[la];
                area += fabs(0.5 * mag3(cross3(A - O, fap - O))) * line_fraction(phi[lo], phi[la]);
            }
            eps_f = area / mag3(Sf);
        }
        volume += (1.0 / 3.0) * eps_f * fabs(dot3(apex - ld3(m.Cf, f), Sf));
    }
    return volume;
}

// ------------------------------------------------------------------------------------------------
// conservative pre-classification of (cell, solid): 0 = no vertex can be inside, 1 = every vertex is
// certainly inside, 2 = evaluate the vertices exactly.  Margins (REL_MARGIN) dwarf fp64 rounding, so
// the exact predicate's outcome is never changed — only skipped when it is certain.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int quick_class(const DevSolid &S, D3 cc, float2 rad) {
    const D3 r = cc - D3{S.pos[0], S.pos[1], S.pos[2]};
    const double d2 = dot3(r, r);
    if (S.kind == KIND_3D) {
        // d - rad > r_out  <=>  d^2 > (r_out + rad)^2 ;  d + rad < r_in  <=>  d^2 < (r_in - rad)^2 with r_in > rad
        // (the 1e-6 relative margins inside r_out / r_in / rad dwarf the rounding of the squares)
        const double ro = S.r_out + (double)rad.x, ri = S.r_in - (double)rad.x;
        if (d2 > ro * ro) return 0;
        if (ri > 0.0 && d2 < ri * ri) return 1;
        return 2;
    }
    if (S.kind == KIND_2D) {
        const double t = r.x * S.axis[0] + r.y * S.axis[1] + r.z * S.axis[2];
        const double dax = sqrt(fmax(0.0, d2 - t * t));
        const double rr = S.axis_is_z ? (double)rad.y : (double)rad.x;
        const double slack = 1e-9 * (sqrt(d2) + 1.0);
        if (dax - rr - slack > S.r_out) return 0;
        if (dax + rr + slack < S.r_in) return 1;
        return 2;
    }
    // plane: body-frame y of the centre
    DQ q = {S.q[0], {S.q[1], S.q[2], S.q[3]}};
    const double yl = world2local(q, D3{S.pos[0], S.pos[1], S.pos[2]}, cc).y;
    const double m = (double)rad.x + 1e-11 * (sqrt(d2) + 1.0);
    if (yl > m) return 0;
    if (yl < -m) return 1;
    return 2;
}

// ------------------------------------------------------------------------------------------------
// K2  interact: parameters shared by the classify / heavy / accumulate kernels
// ------------------------------------------------------------------------------------------------
struct InteractParams {
    DevMesh m;
    const DevSolid *solids;
    const DevShape *shapes;
    int n_solids;
    BinGrid grid;
    const int *bin_off;
    const int *bin_list;
    const int *global_list;
    const double *U;
    double dtINV, rhof;
    double *As, *Fs, *Ts, *Ct;
    double *force_torque; // [6*n_solids], zeroed
    unsigned *pair_counts; // [3*n_solids], zeroed
    int *slots;            // [n_cells*K]: (solid<<2 | type) of every member pair of the cell, -1 terminated
    double *vols;          // [n_cells*K]: solid volume of the pair's cell (boundary types)
    unsigned char *n_item; // [n_cells]: slots in use after k_classify
    int2 *heavy;           // queue of (cell, slot) needing exact evaluation
    unsigned long long *heavy_count;
    long long heavy_cap;
    int n_global;
    int K;
    const unsigned char *excluded; // replay mode: [n_cells*K] 1 = pair is outside the seed's component
    StepStatus *status;
};

#define TPB 128

// warp-level aggregation of one member pair per lane: lanes with the same solid are reduced with a
// butterfly and the group leader issues the 6 fp64 + 1 counter reductions.
__device__ __forceinline__ void warp_accumulate(bool have, int s, int type, const double v[6], double *force_torque,
                                                unsigned *pair_counts) {
    const unsigned FULL = 0xffffffffu;
    unsigned pending = __ballot_sync(FULL, have);
    const int lane = threadIdx.x & 31;
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const int s0 = __shfl_sync(FULL, s, leader);
        const bool mine = have && (s == s0);
        const unsigned grp = __ballot_sync(FULL, mine);
        double w[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) w[k] = mine ? v[k] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 6; ++k) w[k] += __shfl_xor_sync(FULL, w[k], o);
        }
        const unsigned c1 = __popc(__ballot_sync(FULL, mine && type == 1));
        const unsigned c2 = __popc(__ballot_sync(FULL, mine && type == 2));
        const unsigned c3 = __popc(__ballot_sync(FULL, mine && type == 3));
        if (lane == leader) {
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(force_torque + 6 * (long long)s0 + k, w[k]);
            if (c1) atomicAdd(pair_counts + 3 * (long long)s0 + 0, c1);
            if (c2) atomicAdd(pair_counts + 3 * (long long)s0 + 1, c2);
            if (c3) atomicAdd(pair_counts + 3 * (long long)s0 + 2, c3);
        }
        pending &= ~grp;
    }
}

// Exact evaluation of one (cell, solid) item: number of vertices inside, cell type, solid volume.
// General polyhedra: cell-local arrays in local memory, CSR connectivity.
__device__ __noinline__ void heavy_eval_general(const InteractParams &P, int c, int s, int &type_out, double &vol_out) {
    const DevMesh &m = P.m;
    const DevSolid &S = P.solids[s];
    const DevShape &sh = P.shapes[S.shape];
    const DQ q = {S.q[0], {S.q[1], S.q[2], S.q[3]}};
    const D3 t = {S.pos[0], S.pos[1], S.pos[2]};
    int vid[MAX_CELL_VERTS];
    D3 pts[MAX_CELL_VERTS];
    double phi[MAX_CELL_VERTS];
    const int pb = __ldg(m.cp_off + c);
    int nv = __ldg(m.cp_off + c + 1) - pb;
    if (nv > MAX_CELL_VERTS) { nv = MAX_CELL_VERTS; P.status->bad_cell = 1; }
    int n_in = 0;
    for (int k = 0; k < nv; ++k) {
        vid[k] = __ldg(m.cp + pb + k);
        pts[k] = ld3(m.points, vid[k]);
        double ph;
        n_in += shape_eval<true>(sh.s, world2local(q, t, pts[k]), ph) ? 1 : 0;
        phi[k] = ph;
    }
    type_out = 0;
    vol_out = 0.0;
    if (n_in == 0) return;
    if (n_in == nv) { type_out = SDFIBM_CELL_ALL_INSIDE; return; }
    double dummy;
    type_out = shape_eval<false>(sh.s, world2local(q, t, ld3(m.cc, c)), dummy) ? SDFIBM_CELL_CENTER_INSIDE : SDFIBM_CELL_CENTER_OUTSIDE;
    vol_out = cell_solid_volume(m, c, vid, pts, phi, nv);
}

// Hexahedral fast path: fixed 8/6/4 strides, per-cell face->vertex-slot nibbles precomputed at upload,
// vertex coordinates and phi staged in transposed shared memory (column = executing thread) so that the
// data-dependent slot indexing is bank-conflict free.
__device__ __forceinline__ void heavy_eval_hex(const InteractParams &P, int c, int s, int col, double *s_px, double *s_py,
                                               double *s_pz, double *s_phi, int &type_out, double &vol_out) {
    const DevMesh &m = P.m;
    const DevSolid &S = P.solids[s];
    const DevShape &sh = P.shapes[S.shape];
    const DQ q = {S.q[0], {S.q[1], S.q[2], S.q[3]}};
    const D3 t = {S.pos[0], S.pos[1], S.pos[2]};
    const int4 *cp4 = reinterpret_cast<const int4 *>(m.cp + 8 * (long long)c);
    const int4 va = __ldg(cp4), vb = __ldg(cp4 + 1);
    const int vid[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
    int n_in = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const D3 p = ld3(m.points, vid[k]);
        double ph;
        n_in += shape_eval<true>(sh.s, world2local(q, t, p), ph) ? 1 : 0;
        s_px[k * TPB + col] = p.x;
        s_py[k * TPB + col] = p.y;
        s_pz[k * TPB + col] = p.z;
        s_phi[k * TPB + col] = ph;
    }
    type_out = 0;
    vol_out = 0.0;
    if (n_in == 0) return;
    if (n_in == 8) { type_out = SDFIBM_CELL_ALL_INSIDE; return; }
    double dummy;
    type_out = shape_eval<false>(sh.s, world2local(q, t, ld3(m.cc, c)), dummy) ? SDFIBM_CELL_CENTER_INSIDE : SDFIBM_CELL_CENTER_OUTSIDE;

    auto PT = [&](int l) { return D3{s_px[l * TPB + col], s_py[l * TPB + col], s_pz[l * TPB + col]}; };
    auto PH = [&](int l) { return s_phi[l * TPB + col]; };
    // cell apex over the cell's vertex list (geometrictools.cpp:25-45,56-58)
    D3 apex;
    {
        const D3 A = PT(0);
        const double phiA = PH(0);
        D3 B = {0.0, 0.0, 0.0};
        double phiB = 0.0;
        for (int i = 1; i < 8; ++i) {
            B = PT(i);
            phiB = PH(i);
            if (phiA * phiB <= 0) break;
        }
        apex = A - fabs(phiA) / (SDF_SMALL + fabs(phiA) + fabs(phiB)) * (A - B);
        if (m.two_d) apex.z = 0.0;
    }
    const unsigned tw0 = __ldg(m.hex_topo + 3 * (long long)c), tw1 = __ldg(m.hex_topo + 3 * (long long)c + 1),
                   tw2 = __ldg(m.hex_topo + 3 * (long long)c + 2);
    const int2 *cf2 = reinterpret_cast<const int2 *>(m.cf + 6 * (long long)c);
    const int2 f01 = __ldg(cf2), f23 = __ldg(cf2 + 1), f45 = __ldg(cf2 + 2);
    double volume = 0.0;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
        const int face = (f == 0) ? f01.x : (f == 1) ? f01.y : (f == 2) ? f23.x : (f == 3) ? f23.y : (f == 4) ? f45.x : f45.y;
        const unsigned w = (f < 2) ? tw0 : (f < 4) ? tw1 : tw2;
        const unsigned nib = (w >> (16 * (f & 1))) & 0xffffu;
        const int l[4] = {(int)(nib & 0xf), (int)((nib >> 4) & 0xf), (int)((nib >> 8) & 0xf), (int)((nib >> 12) & 0xf)};
        const double ph[4] = {PH(l[0]), PH(l[1]), PH(l[2]), PH(l[3])};
        const int npos = (ph[0] > 0) + (ph[1] > 0) + (ph[2] > 0) + (ph[3] > 0);
        if (npos == 4) continue;                                                // eps_f = 0: adds +0.0 (:107-108)
        double eps_f = 1.0;                                                     // all phi <= 0 (:109-110)
        if (npos != 0) {
            const D3 A = PT(l[0]);                                              // calcFaceArea (:74-96)
            D3 B = PT(l[1]);
            double phiB = ph[1];
            if (!(ph[0] * ph[1] <= 0)) {
                B = PT(l[2]);
                phiB = ph[2];
                if (!(ph[0] * ph[2] <= 0)) { B = PT(l[3]); phiB = ph[3]; }
            }
            const D3 fap = A - fabs(ph[0]) / (SDF_SMALL + fabs(ph[0]) + fabs(phiB)) * (A - B);
            double area = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double lf = line_fraction(ph[e], ph[(e + 1) & 3]);
                if (lf != 0.0) {                                                // a zero fraction adds +0.0
                    const D3 O = PT(l[e]), A2 = PT(l[(e + 1) & 3]);
                    area += fabs(0.5 * mag3(cross3(A2 - O, fap - O))) * lf;
                }
            }
            eps_f = area / __ldg(m.magSf + face);
        }
        volume += (1.0 / 3.0) * eps_f * fabs(dot3(apex - ld3(m.Cf, face), ld3(m.Sf, face)));
    }
    vol_out = volume;
}

// ------------------------------------------------------------------------------------------------
// K2a  k_classify (thread per cell): walk the cell's candidate solids in ascending id; every candidate that
//      may be a member becomes a slot of the cell (certain ALL_INSIDE -> final; otherwise "heavy": appended
//      to a global queue for exact evaluation).
// K2b  k_heavy (thread per heavy item, persistent grid): exact vertex predicates + SDF, cell type,
//      apex/pyramid volume.  Dense: no barriers, no idle lanes waiting on light cells.
// K2c  k_accumulate (thread per cell): consume the cell's slots in ascending solid order — As/Fs/Ts/Ct
//      accumulation exactly in the reference's += order, every field written once and coalesced, per-solid
//      force/torque warp-aggregated before the atomics.
// ------------------------------------------------------------------------------------------------
template <bool HEX>
__global__ void __launch_bounds__(TPB) k_heavy(InteractParams P) {
    __shared__ double s_px[HEX ? 8 * TPB : 1], s_py[HEX ? 8 * TPB : 1], s_pz[HEX ? 8 * TPB : 1], s_phi[HEX ? 8 * TPB : 1];
    const int tid = threadIdx.x;
    const long long n = min((long long)*P.heavy_count, (long long)P.heavy_cap);
    for (long long k = (long long)blockIdx.x * TPB + tid; k < n; k += (long long)gridDim.x * TPB) {
        const int2 it = __ldg(P.heavy + k);           // (cell, slot index)
        const int c = it.x;
        const long long si = (long long)c * P.K + it.y;
        const int s = P.slots[si] >> 2;
        int type;
        double v;
        if (HEX) heavy_eval_hex(P, c, s, tid, s_px, s_py, s_pz, s_phi, type, v);
        else heavy_eval_general(P, c, s, type, v);
        P.slots[si] = (s << 2) | type;                // type 0: no vertex inside -> not a member
        P.vols[si] = v;
    }
}

__global__ void __launch_bounds__(256) k_classify(InteractParams P) {
    const DevMesh &m = P.m;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < m.n_cells;
    int n_item = 0, n_heavy = 0;
    unsigned heavy_mask = 0;
    if (live) {
        const D3 cc = ld3(m.cc, c);
        const float2 rad = m.rad_uniform ? m.rad_const : __ldg(m.cell_rad + c);
        const int b = (bin_coord(P.grid, cc.z, 2) * P.grid.n[1] + bin_coord(P.grid, cc.y, 1)) * P.grid.n[0] +
                      bin_coord(P.grid, cc.x, 0);
        int bi = __ldg(P.bin_off + b);
        const int be = __ldg(P.bin_off + b + 1);
        int gi = 0;
        const int ge = P.status->n_global;
        while (bi < be || gi < ge) {
            // merge the bin list and the global list in ascending solid id
            int s;
            const int sb = (bi < be) ? __ldg(P.bin_list + bi) : 0x7fffffff;
            const int sg = (gi < ge) ? __ldg(P.global_list + gi) : 0x7fffffff;
            if (sb <= sg) { s = sb; ++bi; if (sb == sg) ++gi; }
            else { s = sg; ++gi; }
            const int qc = quick_class(P.solids[s], cc, rad);
            if (qc == 0) continue;
            if (n_item < P.K) {
                P.slots[(long long)c * P.K + n_item] = (s << 2) | qc;   // 1 = certain ALL_INSIDE, 2 = heavy (pending)
                if (qc == 2) { heavy_mask |= 1u << n_item; ++n_heavy; }
                ++n_item;
            } else P.status->slot_overflow = 1;
        }
        P.n_item[c] = (unsigned char)n_item;
    }
    // warp-aggregated append of the heavy items to the global queue
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int incl = n_heavy;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;
    unsigned long long base = 0;
    if (lane == 31) base = atomicAdd(P.heavy_count, (unsigned long long)total);
    base = __shfl_sync(FULL, base, 31);
    long long pos = (long long)base + incl - n_heavy;
    while (heavy_mask) {
        const int j = __ffs(heavy_mask) - 1;
        heavy_mask &= heavy_mask - 1;
        if (pos < P.heavy_cap) P.heavy[pos] = make_int2(c, j);
        ++pos;
    }
}

__global__ void __launch_bounds__(256) k_accumulate(InteractParams P) {
    const DevMesh &m = P.m;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < m.n_cells;
    const int n = live ? (int)P.n_item[c] : 0;
    const unsigned FULL = 0xffffffffu;
    int nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));

    double as = 0.0, ts = 0.0, ct = 0.0;
    D3 fs = {0.0, 0.0, 0.0};
    if (nmax > 0) {
        D3 cc = {0, 0, 0}, uf = {0, 0, 0};
        double vol = 1.0;
        if (n > 0) {
            cc = ld3(m.cc, c);
            uf = ld3(P.U, c);
            vol = __ldg(m.V + c);
        }
        int nslot = 0;
        bool ended = false;
        for (int j = 0; j < nmax; ++j) {
            bool have = false;
            int s = -1, type = 0;
            double contrib[6] = {0, 0, 0, 0, 0, 0};
            if (j < n && !ended) {
                const long long si = (long long)c * P.K + j;
                const int e = P.slots[si];
                if (e < 0) ended = true;                                   // terminator of an earlier (compacting) pass
                else if ((e & 3) != 0) {
                    type = e & 3;
                    s = e >> 2;
                    const double v = (type == SDFIBM_CELL_ALL_INSIDE) ? 0.0 : P.vols[si];
                    const long long so = (long long)c * P.K + nslot;       // compact the members to the front
                    if (nslot != j) { P.slots[so] = e; P.vols[so] = v; }
                    const bool skip = P.excluded && P.excluded[so];        // replay: outside the seed's component
                    ++nslot;
                    if (!skip) {
                        const DevSolid &S = P.solids[s];
                        const D3 t = {S.pos[0], S.pos[1], S.pos[2]};
                        const double alpha = (type == SDFIBM_CELL_ALL_INSIDE) ? 1.0 : v / vol;   // solidcloud.cpp:408-410
                        // solidcloud.cpp:384-390,411-421
                        const D3 om = {S.omega[0], S.omega[1], S.omega[2]};
                        const D3 us = D3{S.vel[0], S.vel[1], S.vel[2]} + cross3(om, cc - t);
                        const D3 f_ = alpha * (uf - us);
                        const D3 t_ = cross3(cc - t, f_);
                        const D3 fo = f_ * vol * P.dtINV;
                        const D3 to = t_ * vol * P.dtINV;
                        contrib[0] = fo.x; contrib[1] = fo.y; contrib[2] = fo.z;
                        contrib[3] = to.x; contrib[4] = to.y; contrib[5] = to.z;
                        as += alpha;
                        fs = fs + f_ * P.dtINV;
                        ts += alpha;
                        ct = (type == SDFIBM_CELL_ALL_INSIDE) ? (double)(s + 4) : (double)type;   // :376-382, last writer wins
                        have = true;
                    }
                }
            }
            if (__any_sync(FULL, have)) warp_accumulate(have, s, type, contrib, P.force_torque, P.pair_counts);
        }
        if (n > 0 && nslot < P.K) P.slots[(long long)c * P.K + nslot] = -1;
        if (live && nslot == 0) ct = 0.0;
        if (live && nslot > 0 && ct == 0.0) ct = -1.0;   // only excluded pairs (replay): slots stay valid for the list extraction
    }
    if (live) {
        P.As[c] = (as < 1.0) ? as : 1.0;                                           // checkAlpha, :564-570 (std::min(As,1))
        P.Fs[3 * (long long)c] = fs.x;
        P.Fs[3 * (long long)c + 1] = fs.y;
        P.Fs[3 * (long long)c + 2] = fs.z;
        P.Ts[c] = ts;
        P.Ct[c] = ct;
    }
}


// replay mode leaves Ct = -1 on cells whose only pairs were excluded; they are untouched cells.
__global__ void k_fix_ct(double *Ct, int n) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && Ct[c] < 0.0) Ct[c] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// K3  connectivity certificate: a member pair is a "root" when no face neighbour that is a member of
// the same solid has a smaller (distance-to-centre, cell id) key.  Exactly one root  =>  the solid's
// vertex-inside cell set is face connected  =>  it equals the reference's flood fill from any seed.
// ------------------------------------------------------------------------------------------------
struct ConnParams {
    DevMesh m;
    const DevSolid *solids;
    const double *Ct;
    const int *slots;
    int K;
    int *root_count; // [n_solids] zeroed
};

__device__ __forceinline__ bool key_less(double ka, int ca, double kb, int cb) { return ka < kb || (ka == kb && ca < cb); }

__global__ void k_connectivity(ConnParams P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.m.n_cells) return;
    if (P.Ct[c] == 0.0) return;
    const D3 cc = ld3(P.m.cc, c);
    const int nb0 = __ldg(P.m.nb_off + c), nb1 = __ldg(P.m.nb_off + c + 1);
    for (int j = 0; j < P.K; ++j) {
        const int e = P.slots[(long long)c * P.K + j];
        if (e < 0) break;
        const int s = e >> 2;
        const D3 x = {P.solids[s].pos[0], P.solids[s].pos[1], P.solids[s].pos[2]};
        const double kc = magSqr3(cc - x);
        bool has_parent = false;
        for (int k = nb0; k < nb1 && !has_parent; ++k) {
            const int nb = __ldg(P.m.nb + k);
            if (P.Ct[nb] == 0.0) continue;
            for (int jj = 0; jj < P.K; ++jj) {
                const int e2 = P.slots[(long long)nb * P.K + jj];
                if (e2 < 0) break;
                if ((e2 >> 2) == s) {
                    const double kn = magSqr3(ld3(P.m.cc, nb) - x);
                    if (key_less(kn, nb, kc, c)) has_parent = true;
                    break;
                }
            }
        }
        if (!has_parent) atomicAdd(P.root_count + s, 1);
    }
}

__global__ void k_finalize(const unsigned *pair_counts, const int *root_count, int n_solids, StepStatus *status) {
    unsigned long long c0 = 0, c1 = 0, c2 = 0;
    int nf = 0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_solids; s += gridDim.x * blockDim.x) {
        c0 += pair_counts[3 * s];
        c1 += pair_counts[3 * s + 1];
        c2 += pair_counts[3 * s + 2];
        nf += root_count[s] > 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        nf += __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c0) atomicAdd(&status->counts[0], c0);
        if (c1) atomicAdd(&status->counts[1], c1);
        if (c2) atomicAdd(&status->counts[2], c2);
        if (nf) atomicAdd(&status->n_flagged, nf);
    }
}

// rhof scaling of the per-solid sums (solidcloud.cpp:424-425)
__global__ void k_scale_ft(double *ft, int n, double rhof) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ft[i] = ft[i] * rhof;
}

// ------------------------------------------------------------------------------------------------
// exact flood-fill replay for solids that failed the certificate (rare path, all on the GPU)
// ------------------------------------------------------------------------------------------------
struct ReplayParams {
    DevMesh m;
    const DevSolid *solids;
    const double *Ct;
    const int *slots;
    int K;
    const int *root_count;
    int *labels;                    // [n_cells*K] component label (min cell id) of flagged pairs
    int *changed;
    unsigned long long *seed_key;   // [n_solids] min dist^2 bits over candidate cells
    int *seed_cell;                 // [n_solids]
    int *min_label;                 // [n_solids]
    int *chosen;                    // [n_solids]
    unsigned char *excluded;        // [n_cells*K]
    BinGrid grid;
    const int *bin_off, *bin_list, *global_list;
    int n_global, n_solids;
};

__global__ void k_replay_init(ReplayParams P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.m.n_cells) return;
    for (int j = 0; j < P.K; ++j) { P.labels[(long long)c * P.K + j] = -1; P.excluded[(long long)c * P.K + j] = 0; }
    if (P.Ct[c] == 0.0) return;
    for (int j = 0; j < P.K; ++j) {
        const int e = P.slots[(long long)c * P.K + j];
        if (e < 0) break;
        if (P.root_count[e >> 2] > 1) P.labels[(long long)c * P.K + j] = c;
    }
}

__global__ void k_replay_propagate(ReplayParams P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.m.n_cells || P.Ct[c] == 0.0) return;
    const int nb0 = __ldg(P.m.nb_off + c), nb1 = __ldg(P.m.nb_off + c + 1);
    for (int j = 0; j < P.K; ++j) {
        const int e = P.slots[(long long)c * P.K + j];
        if (e < 0) break;
        int lab = P.labels[(long long)c * P.K + j];
        if (lab < 0) continue;
        const int s = e >> 2;
        int best = lab;
        for (int k = nb0; k < nb1; ++k) {
            const int nb = __ldg(P.m.nb + k);
            if (P.Ct[nb] == 0.0) continue;
            for (int jj = 0; jj < P.K; ++jj) {
                const int e2 = P.slots[(long long)nb * P.K + jj];
                if (e2 < 0) break;
                if ((e2 >> 2) == s) {
                    const int l2 = ((volatile int *)P.labels)[(long long)nb * P.K + jj];
                    if (l2 >= 0 && l2 < best) best = l2;
                    break;
                }
            }
        }
        if (best < lab) { P.labels[(long long)c * P.K + j] = best; *P.changed = 1; }
    }
}

// nearest cell centre to each flagged solid's centre, restricted to the cells that list the solid as
// a candidate (sufficient: any member cell is within the binned bounding volume).
__global__ void k_replay_seed(ReplayParams P, int pass) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.m.n_cells) return;
    const D3 cc = ld3(P.m.cc, c);
    const int b = (bin_coord(P.grid, cc.z, 2) * P.grid.n[1] + bin_coord(P.grid, cc.y, 1)) * P.grid.n[0] + bin_coord(P.grid, cc.x, 0);
    const int b0 = P.bin_off[b], b1 = P.bin_off[b + 1];
    for (int t = 0; t < (b1 - b0) + P.n_global; ++t) {
        const int s = (t < b1 - b0) ? P.bin_list[b0 + t] : P.global_list[t - (b1 - b0)];
        if (P.root_count[s] <= 1) continue;
        const D3 x = {P.solids[s].pos[0], P.solids[s].pos[1], P.solids[s].pos[2]};
        const unsigned long long key = (unsigned long long)__double_as_longlong(magSqr3(cc - x));
        if (pass == 0) atomicMin(P.seed_key + s, key);
        else if (key == P.seed_key[s]) atomicMin(P.seed_cell + s, c);
    }
    if (pass == 1 && P.Ct[c] != 0.0) {
        for (int j = 0; j < P.K; ++j) {
            const int e = P.slots[(long long)c * P.K + j];
            if (e < 0) break;
            const int lab = P.labels[(long long)c * P.K + j];
            if (lab >= 0) atomicMin(P.min_label + (e >> 2), lab);
        }
    }
}

__global__ void k_replay_choose(ReplayParams P) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_solids) return;
    if (P.root_count[s] <= 1) { P.chosen[s] = -1; return; }
    int chosen = P.min_label[s]; // component of the first member cell in index order (cellenumerator.cpp:52-63)
    const int g = P.seed_cell[s];
    if (g >= 0 && g < P.m.n_cells && P.Ct[g] != 0.0) {
        for (int j = 0; j < P.K; ++j) {
            const int e = P.slots[(long long)g * P.K + j];
            if (e < 0) break;
            if ((e >> 2) == s) { chosen = P.labels[(long long)g * P.K + j]; break; } // nearest cell is a member: it is the seed
        }
    }
    P.chosen[s] = chosen;
}

__global__ void k_replay_mark(ReplayParams P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.m.n_cells || P.Ct[c] == 0.0) return;
    for (int j = 0; j < P.K; ++j) {
        const int e = P.slots[(long long)c * P.K + j];
        if (e < 0) break;
        const int lab = P.labels[(long long)c * P.K + j];
        if (lab >= 0 && lab != P.chosen[e >> 2]) P.excluded[(long long)c * P.K + j] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// K4  fixInternal (solidcloud.cpp:288-301)
// ------------------------------------------------------------------------------------------------
__global__ void k_fix_internal(DevMesh m, const sdfibm_solid_t *solids, int n_solids, const double *Ct, double *U) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.n_cells) return;
    const double ct = Ct[c];
    if (ct >= 4) {
        const int id = (int)(ct - 4);
        if (id < n_solids) {
            const sdfibm_solid_t &S = solids[id];
            const D3 x = {S.pos[0], S.pos[1], S.pos[2]};
            const D3 u = D3{S.vel[0], S.vel[1], S.vel[2]} + cross3(D3{S.omega[0], S.omega[1], S.omega[2]}, ld3(m.cc, c) - x);
            U[3 * (long long)c] = u.x;
            U[3 * (long long)c + 1] = u.y;
            U[3 * (long long)c + 2] = u.z;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// candidate list extraction (parity output, off the timed path): pairs in cell order, then a stable
// radix sort by (solid, type) gives ascending cell ids inside every segment (std::set order).
// ------------------------------------------------------------------------------------------------
__global__ void k_list_count(const double *Ct, const int *slots, const unsigned char *excluded, int K, int n_cells, int *cnt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    int n = 0;
    if (Ct[c] != 0.0)
        for (int j = 0; j < K; ++j) {
            if (slots[(long long)c * K + j] < 0) break;
            if (!(excluded && excluded[(long long)c * K + j])) ++n;
        }
    cnt[c] = n;
}
__global__ void k_list_emit(const double *Ct, const int *slots, const unsigned char *excluded, int K, int n_cells,
                            const int *off, unsigned *keys, int *vals) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells || Ct[c] == 0.0) return;
    int o = off[c];
    for (int j = 0; j < K; ++j) {
        const int e = slots[(long long)c * K + j];
        if (e < 0) break;
        if (excluded && excluded[(long long)c * K + j]) continue;
        keys[o] = (unsigned)(3 * (e >> 2) + ((e & 3) - 1));
        vals[o] = c;
        ++o;
    }
}

// ------------------------------------------------------------------------------------------------
// collision step (solidcloud.cpp:477-519, libcollision/): centres hashed on the UGrid, pairs
// enumerated per (grid cell, neighbour cell) in the reference's i,j,k / 27-neighbour order, one
// thread per contact pair for the narrow phase and force law.
// ------------------------------------------------------------------------------------------------
struct UGridDev {
    double lo[3];
    double deltaINV;
    int nx, ny, nz, nynz;
};
__device__ __forceinline__ int ugrid_hash(const UGridDev &g, const double *pos) {
    int i = (int)floor((pos[0] - g.lo[0]) * g.deltaINV);
    int j = (int)floor((pos[1] - g.lo[1]) * g.deltaINV);
    int k = (int)floor((pos[2] - g.lo[2]) * g.deltaINV);
    return i * g.nynz + j * g.nz + k;
}
// keys are stored order-preserving as unsigned (signed hash ^ 0x80000000) for the radix sort
__device__ __forceinline__ unsigned ukey_of(int key) { return (unsigned)key ^ 0x80000000u; }
__global__ void k_col_keys(const sdfibm_solid_t *solids, int n, UGridDev g, unsigned *keys, int *ids) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    keys[s] = ukey_of(ugrid_hash(g, solids[s].pos));
    ids[s] = s;
}
// skeys ascending (stable: ids ascending inside a key == push_back order, ugrid.h:29-32).  The thread of
// the FIRST solid of every occupied grid cell enumerates that cell's pairs in the reference's order
// (ugrid.cpp:52-74): neighbour cell (27, i/j/k nested) -> pi in own list -> qi in neighbour list, pi < qi.
__device__ __forceinline__ int lower_bound_key(const unsigned *keys, int n, unsigned key) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
__global__ void k_col_pairs(const unsigned *skeys, const int *sids, int n, UGridDev g, int *pair_cnt, const int *pair_off,
                            int *pairs, long long cap, int emit) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned uk = skeys[t];
    if (t > 0 && skeys[t - 1] == uk) { if (!emit) pair_cnt[t] = 0; return; }
    const int key = (int)(uk ^ 0x80000000u);
    int cnt = 0;
    long long o = emit ? pair_off[t] : 0;
    // out-of-range keys never own a loop iteration in the reference; aliased ones behave as the aliased cell
    if (key >= 0 && key < g.nx * g.nynz) {
        int own_end = t;
        while (own_end < n && skeys[own_end] == uk) ++own_end;
        const int i = key / g.nynz, j = (key - i * g.nynz) / g.nz, k = key - i * g.nynz - j * g.nz;
        for (int nbi = i - 1; nbi <= i + 1; ++nbi)
            for (int nbj = j - 1; nbj <= j + 1; ++nbj)
                for (int nbk = k - 1; nbk <= k + 1; ++nbk) {
                    if (nbi < 0 || nbi > g.nx - 1 || nbj < 0 || nbj > g.ny - 1 || nbk < 0 || nbk > g.nz - 1) continue;
                    const unsigned nkey = ukey_of(nbi * g.nynz + nbj * g.nz + nbk);
                    const int u0 = lower_bound_key(skeys, n, nkey);
                    for (int a = t; a < own_end; ++a) {
                        const int p = sids[a];
                        for (int u = u0; u < n && skeys[u] == nkey; ++u) {
                            const int qv = sids[u];
                            if (p < qv) {
                                if (emit && o < cap) { pairs[2 * o] = p; pairs[2 * o + 1] = qv; }
                                ++o;
                                ++cnt;
                            }
                        }
                    }
                }
    }
    if (!emit) pair_cnt[t] = cnt;
}
__global__ void k_col_narrow(const sdfibm_solid_t *solids, const DevShape *shapes, const int *pairs, long long n_pairs,
                             double *ft) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const int i1 = pairs[2 * t], i2 = pairs[2 * t + 1];
    const sdfibm_solid_t &s1 = solids[i1], &s2 = solids[i2];
    // SHAPE2ID with operator[] default 0 (collision.h:11-15): Plane 0, Circle 1, Sphere 2, everything else 0
    int a = shapes[s1.shape].s.tag, b = shapes[s2.shape].s.tag;
    a = (a == SDFIBM_SHAPE_CIRCLE || a == SDFIBM_SHAPE_SPHERE) ? a : 0;
    b = (b == SDFIBM_SHAPE_CIRCLE || b == SDFIBM_SHAPE_SPHERE) ? b : 0;
    double cd;
    D3 cN;
    const D3 c1 = {s1.pos[0], s1.pos[1], s1.pos[2]}, c2 = {s2.pos[0], s2.pos[1], s2.pos[2]};
    if (a == 0 && b == 0) return;
    if (a == 0 || b == 0) {                                                    // planeSphereCollision, collision.cpp:22-30
        const sdfibm_solid_t &p = (a == 0) ? s1 : s2;
        const sdfibm_solid_t &s = (a == 0) ? s2 : s1;
        const DQ q = {p.quat[0], {p.quat[1], p.quat[2], p.quat[3]}};
        const D3 sc = world2local(q, D3{p.pos[0], p.pos[1], p.pos[2]}, D3{s.pos[0], s.pos[1], s.pos[2]});
        cN = qtransform(q, D3{0.0, 1.0, 0.0});
        cd = shapes[s.shape].s.radiusB - sc.y;
    } else if (a == b) {                                                       // sphereSphereCollision, :7-18
        const D3 s2s = c2 - c1;
        const double ms = mag3(s2s);
        cN = (ms > 1e-300) ? s2s / ms : D3{0.0, 0.0, 0.0};
        cd = shapes[s1.shape].s.radiusB + shapes[s2.shape].s.radiusB - ms;
    } else return;
    if (cd < 0) return;                                                        // solidcloud.cpp:509-510
    const D3 force = (1e4 * cd) * cN;                                          // :511
    atomicAdd(ft + 6 * (long long)i1 + 0, -force.x);
    atomicAdd(ft + 6 * (long long)i1 + 1, -force.y);
    atomicAdd(ft + 6 * (long long)i1 + 2, -force.z);
    atomicAdd(ft + 6 * (long long)i2 + 0, force.x);
    atomicAdd(ft + 6 * (long long)i2 + 1, force.y);
    atomicAdd(ft + 6 * (long long)i2 + 2, force.z);
}

// =================================================================================================
// host side: context
// =================================================================================================
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t want) {
        if (want <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(want, 1) * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct sdfibm_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    int K = 4;
    // mesh
    bool has_mesh = false;
    DevMesh dm{};
    DevBuf<double> points, cc, V, Cf, Sf;
    DevBuf<int> cp_off, cp, cf_off, cf, fp_off, fp, nb_off, nb;
    DevBuf<float2> cell_rad;
    DevBuf<double> magSf;
    DevBuf<unsigned> hex_topo;
    double bmin[3], bmax[3];
    float rad3_max = 0.f, radxy_max = 0.f;
    // shapes
    std::vector<DevShape> h_shapes;
    DevBuf<DevShape> shapes;
    // per step
    DevBuf<sdfibm_solid_t> solids_in;
    DevBuf<DevSolid> solids;
    DevBuf<int> bin_count, bin_off, bin_cursor, bin_list, global_list, root_count, slots;
    DevBuf<double> vols;
    DevBuf<unsigned char> n_item;
    DevBuf<int2> heavy;
    DevBuf<unsigned> pair_counts;
    DevBuf<double> ft_internal;
    DevBuf<StepStatus> status;
    DevBuf<unsigned char> scan_tmp;
    StepStatus *h_status = nullptr; // pinned
    sdfibm_solid_t *h_solids = nullptr; // pinned staging
    size_t h_solids_cap = 0;
    BinGrid grid{};
    int n_solids_last = 0;
    // fields kept on device for the host-buffer API and fixInternal
    DevBuf<double> dU, dAs, dFs, dTs, dCt, dFT;
    const double *last_Ct = nullptr;
    // replay
    DevBuf<int> labels, seed_cell, min_label, chosen, changed;
    DevBuf<unsigned long long> seed_key;
    DevBuf<unsigned char> excluded;
    bool last_used_replay = false;
    // stats
    StepStatus last{};
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double t_ms[6] = {0, 0, 0, 0, 0, 0}; // binning, k_classify, k_heavy, k_accumulate, connectivity+finalise, whole pipeline
    int n_sm = 148;
    int64_t launches = 0;
    int64_t flagged_last = 0;
};

static int grid_for(long long n, int block) { return (int)std::max<long long>(1, (n + block - 1) / block); }

template <typename T>
static int upload(DevBuf<T> &b, const T *src, size_t n, cudaStream_t st) {
    CUDA_TRY(b.ensure(n));
    if (n) CUDA_TRY(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return SDFIBM_OK;
}

static void shape_bounds(const sdfibm_shape_t &s, DevShape &d) {
    const double *p = s.p;
    const double com = std::sqrt(s.com[0] * s.com[0] + s.com[1] * s.com[1] + s.com[2] * s.com[2]);
    double ro = 0, ri = 0;
    int kind = KIND_3D;
    double comv = com;
    switch (s.tag) {
    case SDFIBM_SHAPE_PLANE: kind = KIND_PLANE; break;
    case SDFIBM_SHAPE_CIRCLE: kind = KIND_2D; ro = p[0]; ri = p[0]; break;
    case SDFIBM_SHAPE_SPHERE: ro = p[0]; ri = p[0]; break;
    case SDFIBM_SHAPE_ELLIPSE: kind = KIND_2D; ro = std::max(p[0], p[1]); ri = std::min(p[0], p[1]); break;
    case SDFIBM_SHAPE_ELLIPSOID: ro = std::max(p[0], std::max(p[1], p[2])); ri = std::min(p[0], std::min(p[1], p[2])); comv = 0; break;
    case SDFIBM_SHAPE_RECTANGLE: kind = KIND_2D; ro = std::sqrt(p[0] * p[0] + p[1] * p[1]); ri = std::min(p[0], p[1]); break;
    case SDFIBM_SHAPE_BOX: ro = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]); ri = std::min(p[0], std::min(p[1], p[2])); break;
    case SDFIBM_SHAPE_CIRCLE_TAIL:
    case SDFIBM_SHAPE_CIRCLE_TWOTAIL:
        kind = KIND_2D;
        ro = std::max(p[0], std::sqrt(4 * p[2] * p[2] + p[3] * p[3]) * 1.001); // tail box [0,2A]x[-w,w]; rot30 literal is not unitary
        ri = p[0];
        break;
    }
    d.s = s;
    d.kind = kind;
    d.r_out = (ro + comv) * (1.0 + REL_MARGIN) + 1e-300;
    d.r_in = std::max(0.0, (ri - comv) * (1.0 - REL_MARGIN));
    if (!(d.r_out < 1e300)) d.r_out = 1e300;
    d.pad = 0;
}

extern "C" {

int sdfibm_version(void) { return 100; }
const char *sdfibm_last_error(void) { return g_last_error.c_str(); }

int sdfibm_device_count(int *count) {
    if (!count) return fail(SDFIBM_ERR_ARG, "null count");
    CUDA_TRY(cudaGetDeviceCount(count));
    return SDFIBM_OK;
}

int sdfibm_create(int device, sdfibm_context **out) {
    if (!out) return fail(SDFIBM_ERR_ARG, "sdfibm_create: null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SDFIBM_ERR_CUDA, std::string("sdfibm_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(SDFIBM_ERR_ARG, "sdfibm_create: bad device index");
    CUDA_TRY(cudaSetDevice(device));
    auto *ctx = new sdfibm_context();
    ctx->device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMallocHost(&ctx->h_status, sizeof(StepStatus)));
    CUDA_TRY(ctx->status.ensure(1));
    for (int i = 0; i < 6; ++i) CUDA_TRY(cudaEventCreate(&ctx->ev[i]));
    CUDA_TRY(cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device));
    *out = ctx;
    return SDFIBM_OK;
}

int sdfibm_destroy(sdfibm_context *ctx) {
    if (!ctx) return SDFIBM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->points.release(); ctx->cc.release(); ctx->V.release(); ctx->Cf.release(); ctx->Sf.release();
    ctx->cp_off.release(); ctx->cp.release(); ctx->cf_off.release(); ctx->cf.release();
    ctx->fp_off.release(); ctx->fp.release(); ctx->nb_off.release(); ctx->nb.release();
    ctx->cell_rad.release(); ctx->magSf.release(); ctx->hex_topo.release(); ctx->shapes.release(); ctx->solids_in.release(); ctx->solids.release();
    ctx->bin_count.release(); ctx->bin_off.release(); ctx->bin_cursor.release(); ctx->bin_list.release();
    ctx->global_list.release(); ctx->root_count.release(); ctx->slots.release(); ctx->pair_counts.release();
    ctx->vols.release(); ctx->n_item.release(); ctx->heavy.release();
    ctx->ft_internal.release(); ctx->status.release(); ctx->scan_tmp.release();
    ctx->dU.release(); ctx->dAs.release(); ctx->dFs.release(); ctx->dTs.release(); ctx->dCt.release(); ctx->dFT.release();
    ctx->labels.release(); ctx->seed_cell.release(); ctx->min_label.release(); ctx->chosen.release();
    ctx->changed.release(); ctx->seed_key.release(); ctx->excluded.release();
    for (int i = 0; i < 6; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_solids) cudaFreeHost(ctx->h_solids);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SDFIBM_OK;
}

int sdfibm_set_cell_slots(sdfibm_context *ctx, int slots) {
    if (!ctx || slots < 1 || slots > 64) return fail(SDFIBM_ERR_ARG, "sdfibm_set_cell_slots: slots must be in 1..64");
    if (ctx->has_mesh) return fail(SDFIBM_ERR_STATE, "sdfibm_set_cell_slots: call before sdfibm_set_mesh");
    ctx->K = slots;
    return SDFIBM_OK;
}

int sdfibm_stream(sdfibm_context *ctx, void **s) {
    if (!ctx || !s) return fail(SDFIBM_ERR_ARG, "sdfibm_stream: null argument");
    *s = (void *)ctx->stream;
    return SDFIBM_OK;
}
int sdfibm_synchronize(sdfibm_context *ctx) {
    if (!ctx) return fail(SDFIBM_ERR_ARG, "null context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SDFIBM_OK;
}

int sdfibm_set_mesh(sdfibm_context *ctx, const sdfibm_mesh_t *m, int two_d) {
    if (!ctx || !m) return fail(SDFIBM_ERR_ARG, "sdfibm_set_mesh: null argument");
    if (m->n_cells <= 0 || !m->points || !m->cell_centres || !m->cell_volumes || !m->face_centres || !m->face_areas ||
        !m->cell_points_off || !m->cell_points || !m->cell_faces_off || !m->cell_faces || !m->face_points_off ||
        !m->face_points || !m->cell_cells_off || !m->cell_cells)
        return fail(SDFIBM_ERR_ARG, "sdfibm_set_mesh: missing mesh array");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nC = m->n_cells, nP = m->n_points, nF = m->n_faces;
    int rc;
    if ((rc = upload(ctx->points, m->points, 3 * nP, st))) return rc;
    if ((rc = upload(ctx->cc, m->cell_centres, 3 * nC, st))) return rc;
    if ((rc = upload(ctx->V, m->cell_volumes, nC, st))) return rc;
    if ((rc = upload(ctx->Cf, m->face_centres, 3 * nF, st))) return rc;
    if ((rc = upload(ctx->Sf, m->face_areas, 3 * nF, st))) return rc;
    if ((rc = upload(ctx->cp_off, m->cell_points_off, nC + 1, st))) return rc;
    if ((rc = upload(ctx->cp, m->cell_points, (size_t)m->cell_points_off[nC], st))) return rc;
    if ((rc = upload(ctx->cf_off, m->cell_faces_off, nC + 1, st))) return rc;
    if ((rc = upload(ctx->cf, m->cell_faces, (size_t)m->cell_faces_off[nC], st))) return rc;
    if ((rc = upload(ctx->fp_off, m->face_points_off, nF + 1, st))) return rc;
    if ((rc = upload(ctx->fp, m->face_points, (size_t)m->face_points_off[nF], st))) return rc;
    if ((rc = upload(ctx->nb_off, m->cell_cells_off, nC + 1, st))) return rc;
    if ((rc = upload(ctx->nb, m->cell_cells, (size_t)m->cell_cells_off[nC], st))) return rc;
    CUDA_TRY(ctx->cell_rad.ensure(nC));
    CUDA_TRY(ctx->slots.ensure(nC * ctx->K));
    CUDA_TRY(ctx->vols.ensure(nC * ctx->K));
    CUDA_TRY(ctx->n_item.ensure(nC));
    CUDA_TRY(ctx->heavy.ensure(std::max<size_t>(1 << 20, nC / 2)));
    DevMesh &d = ctx->dm;
    d.n_cells = m->n_cells; d.n_points = m->n_points; d.n_faces = m->n_faces;
    d.points = ctx->points.p; d.cc = ctx->cc.p; d.V = ctx->V.p; d.Cf = ctx->Cf.p; d.Sf = ctx->Sf.p;
    d.cp_off = ctx->cp_off.p; d.cp = ctx->cp.p; d.cf_off = ctx->cf_off.p; d.cf = ctx->cf.p;
    d.fp_off = ctx->fp_off.p; d.fp = ctx->fp.p; d.nb_off = ctx->nb_off.p; d.nb = ctx->nb.p;
    d.cell_rad = ctx->cell_rad.p;
    d.two_d = two_d ? 1 : 0;
    // hexahedral fast path?
    bool is_hex = true;
    for (size_t c = 0; c < nC && is_hex; ++c)
        is_hex = (m->cell_points_off[c + 1] - m->cell_points_off[c] == 8) && (m->cell_faces_off[c + 1] - m->cell_faces_off[c] == 6);
    for (size_t f = 0; f < nF && is_hex; ++f) is_hex = (m->face_points_off[f + 1] - m->face_points_off[f] == 4);
    is_hex = is_hex && m->cell_points_off[0] == 0 && m->cell_faces_off[0] == 0 && m->face_points_off[0] == 0;
    d.is_hex = is_hex ? 1 : 0;
    CUDA_TRY(ctx->magSf.ensure(nF));
    d.magSf = ctx->magSf.p;
    d.hex_topo = nullptr;
    k_face_mag<<<grid_for(nF, 256), 256, 0, st>>>(d.Sf, (int)nF, ctx->magSf.p);
    for (int k = 0; k < 3; ++k) { ctx->bmin[k] = m->bounds_min[k]; ctx->bmax[k] = m->bounds_max[k]; }
    // per-cell radii + maxima
    DevBuf<int> bad;
    DevBuf<float> rmax;
    CUDA_TRY(bad.ensure(1));
    CUDA_TRY(rmax.ensure(4));
    CUDA_TRY(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    {
        const float init[4] = {0.f, 0.f, 3.0e38f, 3.0e38f};
        CUDA_TRY(cudaMemcpyAsync(rmax.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    k_cell_radius<<<grid_for(nC, 256), 256, 0, st>>>(d, ctx->cell_rad.p, bad.p, rmax.p);
    if (is_hex) {
        CUDA_TRY(ctx->hex_topo.ensure(3 * nC));
        d.hex_topo = ctx->hex_topo.p;
        k_hex_topo<<<grid_for(nC, 256), 256, 0, st>>>(d, ctx->hex_topo.p, bad.p);
    }
    CUDA_TRY(cudaGetLastError());
    int h_bad = 0;
    float h_rmax[4];
    CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_rmax, rmax.p, 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    bad.release();
    rmax.release();
    if (h_bad) return fail(SDFIBM_ERR_UNSUPPORTED, "sdfibm_set_mesh: a cell has more than 32 vertices, a face more than 16, or inconsistent cell/face point lists");
    ctx->rad3_max = h_rmax[0];
    ctx->radxy_max = h_rmax[1];
    // near-uniform cell sizes: the mesh-wide upper bound serves every cell (always conservative) and the
    // per-cell radius array is not read by the interact kernel
    d.rad_const = make_float2(h_rmax[0], h_rmax[1]);
    d.rad_uniform = (h_rmax[0] <= 1.05f * h_rmax[2] && h_rmax[1] <= 1.05f * h_rmax[3]) ? 1 : 0;
    ctx->has_mesh = true;
    ctx->last_Ct = nullptr;
    return SDFIBM_OK;
}

int sdfibm_set_shapes(sdfibm_context *ctx, const sdfibm_shape_t *shapes, int n) {
    if (!ctx || !shapes || n <= 0) return fail(SDFIBM_ERR_ARG, "sdfibm_set_shapes: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx->h_shapes.resize(n);
    for (int i = 0; i < n; ++i) {
        if (shapes[i].tag < 0 || shapes[i].tag >= SDFIBM_SHAPE_NTAGS)
            return fail(SDFIBM_ERR_UNSUPPORTED, "sdfibm_set_shapes: shape type has no device tag (no CPU fallback)");
        shape_bounds(shapes[i], ctx->h_shapes[i]);
    }
    int rc = upload(ctx->shapes, ctx->h_shapes.data(), (size_t)n, ctx->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SDFIBM_OK;
}

// choose the solid-binning grid for this step
static void choose_grid(sdfibm_context *ctx) {
    double rmax = 0.0;
    for (auto &s : ctx->h_shapes)
        if (s.kind != KIND_PLANE) rmax = std::max(rmax, s.r_out);
    double ext[3], vol = 1.0;
    int nd = 0;
    for (int d = 0; d < 3; ++d) {
        ext[d] = std::max(ctx->bmax[d] - ctx->bmin[d], 1e-300);
        vol *= ext[d];
    }
    double b = std::max(0.75 * rmax, 1e-300);
    // never more than ~4M bins and never finer than the cells
    double bmin_cells = std::cbrt(vol / std::max(1.0, (double)ctx->dm.n_cells)) * 2.0;
    b = std::max(b, bmin_cells);
    for (;;) {
        double nb = 1;
        for (int d = 0; d < 3; ++d) nb *= std::max(1.0, std::ceil(ext[d] / b));
        if (nb <= 4.0e6) break;
        b *= 1.26;
    }
    (void)nd;
    BinGrid &g = ctx->grid;
    g.inv_b = 1.0 / b;
    g.n_bins = 1;
    for (int d = 0; d < 3; ++d) {
        g.lo[d] = ctx->bmin[d];
        g.n[d] = (int)std::max(1.0, std::ceil(ext[d] / b));
        g.n_bins *= g.n[d];
    }
}

static int run_pipeline(sdfibm_context *ctx, int n_solids, const double *dU, double dt, double rhof, double *dAs,
                        double *dFs, double *dTs, double *dCt, double *dFT, bool replay);

static int stage_solids(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n) {
    if ((size_t)n > ctx->h_solids_cap) {
        if (ctx->h_solids) cudaFreeHost(ctx->h_solids);
        ctx->h_solids = nullptr;
        ctx->h_solids_cap = 0;
        CUDA_TRY(cudaMallocHost(&ctx->h_solids, sizeof(sdfibm_solid_t) * (size_t)n));
        ctx->h_solids_cap = n;
    }
    const int ns = (int)ctx->h_shapes.size();
    for (int i = 0; i < n; ++i)
        if (solids[i].shape < 0 || solids[i].shape >= ns) return fail(SDFIBM_ERR_ARG, "solid refers to an unknown shape index");
    memcpy(ctx->h_solids, solids, sizeof(sdfibm_solid_t) * (size_t)n);
    CUDA_TRY(ctx->solids_in.ensure(n));
    CUDA_TRY(cudaMemcpyAsync(ctx->solids_in.p, ctx->h_solids, sizeof(sdfibm_solid_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    return SDFIBM_OK;
}

int sdfibm_interact_device(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, const double *dU, double dt,
                           double rhof, double *dAs, double *dFs, double *dTs, double *dCt, double *dFT) {
    if (!ctx || !solids || n_solids <= 0 || !dU || !dAs || !dFs || !dTs || !dCt || !dFT)
        return fail(SDFIBM_ERR_ARG, "sdfibm_interact: null argument");
    if (!ctx->has_mesh || ctx->h_shapes.empty()) return fail(SDFIBM_ERR_STATE, "sdfibm_interact: set mesh and shapes first");
    if (n_solids > (1 << 29) - 4) return fail(SDFIBM_ERR_ARG, "too many solids");
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = stage_solids(ctx, solids, n_solids);
    if (rc) return rc;
    ctx->launches = 0;
    rc = run_pipeline(ctx, n_solids, dU, dt, rhof, dAs, dFs, dTs, dCt, dFT, false);
    if (rc) return rc;
    ctx->flagged_last = ctx->last.n_flagged;
    ctx->last_used_replay = false;
    if (ctx->last.n_flagged > 0) {
        // exact flood-fill replay for the flagged solids
        const size_t nC = ctx->dm.n_cells, K = ctx->K;
        cudaStream_t st = ctx->stream;
        CUDA_TRY(ctx->labels.ensure(nC * K));
        CUDA_TRY(ctx->excluded.ensure(nC * K));
        CUDA_TRY(ctx->seed_key.ensure(n_solids));
        CUDA_TRY(ctx->seed_cell.ensure(n_solids));
        CUDA_TRY(ctx->min_label.ensure(n_solids));
        CUDA_TRY(ctx->chosen.ensure(n_solids));
        CUDA_TRY(ctx->changed.ensure(1));
        ReplayParams R;
        R.m = ctx->dm; R.solids = ctx->solids.p; R.Ct = dCt; R.slots = ctx->slots.p; R.K = ctx->K;
        R.root_count = ctx->root_count.p; R.labels = ctx->labels.p; R.changed = ctx->changed.p;
        R.seed_key = ctx->seed_key.p; R.seed_cell = ctx->seed_cell.p; R.min_label = ctx->min_label.p;
        R.chosen = ctx->chosen.p; R.excluded = ctx->excluded.p; R.grid = ctx->grid; R.bin_off = ctx->bin_off.p;
        R.bin_list = ctx->bin_list.p; R.global_list = ctx->global_list.p; R.n_global = ctx->last.n_global;
        R.n_solids = n_solids;
        const int g = grid_for(nC, 256);
        k_replay_init<<<g, 256, 0, st>>>(R);
        CUDA_TRY(cudaMemsetAsync(ctx->seed_key.p, 0xff, sizeof(unsigned long long) * n_solids, st));
        CUDA_TRY(cudaMemsetAsync(ctx->seed_cell.p, 0x7f, sizeof(int) * n_solids, st));
        CUDA_TRY(cudaMemsetAsync(ctx->min_label.p, 0x7f, sizeof(int) * n_solids, st));
        for (int it = 0; it < 1000000; ++it) {
            int h_changed = 0;
            CUDA_TRY(cudaMemsetAsync(ctx->changed.p, 0, sizeof(int), st));
            for (int rep = 0; rep < 8; ++rep) k_replay_propagate<<<g, 256, 0, st>>>(R);
            CUDA_TRY(cudaMemcpyAsync(&h_changed, ctx->changed.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            ctx->launches += 8;
            if (!h_changed) break;
        }
        k_replay_seed<<<g, 256, 0, st>>>(R, 0);
        k_replay_seed<<<g, 256, 0, st>>>(R, 1);
        k_replay_choose<<<grid_for(n_solids, 256), 256, 0, st>>>(R);
        k_replay_mark<<<g, 256, 0, st>>>(R);
        CUDA_TRY(cudaGetLastError());
        ctx->launches += 5;
        rc = run_pipeline(ctx, n_solids, dU, dt, rhof, dAs, dFs, dTs, dCt, dFT, true);
        if (rc) return rc;
        ctx->last_used_replay = true;
    }
    ctx->last_Ct = dCt;
    ctx->n_solids_last = n_solids;
    return SDFIBM_OK;
}

static int run_pipeline(sdfibm_context *ctx, int n_solids, const double *dU, double dt, double rhof, double *dAs,
                        double *dFs, double *dTs, double *dCt, double *dFT, bool replay) {
    cudaStream_t st = ctx->stream;
    const int nC = ctx->dm.n_cells;
    if (!replay) {
        choose_grid(ctx);
        const BinGrid &g = ctx->grid;
        CUDA_TRY(ctx->solids.ensure(n_solids));
        CUDA_TRY(ctx->bin_count.ensure((size_t)g.n_bins + 1));
        CUDA_TRY(ctx->bin_off.ensure((size_t)g.n_bins + 1));
        CUDA_TRY(ctx->bin_cursor.ensure((size_t)g.n_bins));
        CUDA_TRY(ctx->global_list.ensure(n_solids));
        CUDA_TRY(ctx->root_count.ensure(n_solids));
        CUDA_TRY(ctx->pair_counts.ensure(3 * (size_t)n_solids));
        if (ctx->bin_list.n == 0) CUDA_TRY(ctx->bin_list.ensure(std::max<size_t>(1 << 20, 128 * (size_t)n_solids)));
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        const BinGrid &g = ctx->grid;
        CUDA_TRY(cudaEventRecord(ctx->ev[0], st));
        CUDA_TRY(cudaMemsetAsync(ctx->status.p, 0, sizeof(StepStatus), st));
        CUDA_TRY(cudaMemsetAsync(ctx->root_count.p, 0, sizeof(int) * n_solids, st));
        CUDA_TRY(cudaMemsetAsync(ctx->pair_counts.p, 0, sizeof(unsigned) * 3 * n_solids, st));
        CUDA_TRY(cudaMemsetAsync(dFT, 0, sizeof(double) * 6 * n_solids, st));
        if (!replay) {
            CUDA_TRY(cudaMemsetAsync(ctx->bin_count.p, 0, sizeof(int) * ((size_t)g.n_bins + 1), st));
            CUDA_TRY(cudaMemsetAsync(ctx->bin_cursor.p, 0, sizeof(int) * (size_t)g.n_bins, st));
            PrepParams P;
            P.solids = ctx->solids_in.p; P.shapes = ctx->shapes.p; P.n_solids = n_solids; P.n_shapes = (int)ctx->h_shapes.size();
            P.out = ctx->solids.p; P.grid = g; P.rad3_max = ctx->rad3_max; P.radxy_max = ctx->radxy_max;
            for (int d = 0; d < 3; ++d) { P.mesh_lo[d] = ctx->bmin[d]; P.mesh_hi[d] = ctx->bmax[d]; }
            P.bin_count = ctx->bin_count.p; P.global_list = ctx->global_list.p; P.status = ctx->status.p;
            k_solid_prepare<<<grid_for(n_solids, 128), 128, 0, st>>>(P);
            size_t tmp_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->bin_count.p, ctx->bin_off.p, g.n_bins + 1, st);
            CUDA_TRY(ctx->scan_tmp.ensure(tmp_bytes));
            cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp_bytes, ctx->bin_count.p, ctx->bin_off.p, g.n_bins + 1, st);
            FillParams F;
            F.solids = ctx->solids.p; F.n_solids = n_solids; F.grid = g; F.rad3_max = ctx->rad3_max; F.radxy_max = ctx->radxy_max;
            for (int d = 0; d < 3; ++d) { F.mesh_lo[d] = ctx->bmin[d]; F.mesh_hi[d] = ctx->bmax[d]; }
            F.bin_off = ctx->bin_off.p; F.bin_cursor = ctx->bin_cursor.p; F.bin_list = ctx->bin_list.p;
            F.bin_cap = (int)std::min<size_t>(ctx->bin_list.n, 0x7fffffff); F.status = ctx->status.p;
            k_bin_fill<<<grid_for(n_solids, 128), 128, 0, st>>>(F);
            k_bin_sort<<<grid_for((long long)g.n_bins + 1, 128), 128, 0, st>>>(ctx->bin_off.p, ctx->bin_list.p, g.n_bins, F.bin_cap,
                                                                              ctx->global_list.p, ctx->status.p);
            ctx->launches += 4;
        } else {
            // keep the binning of the first pass; restore the counters the status word carries
            StepStatus keep{};
            keep.n_global = ctx->last.n_global;
            keep.bin_total = ctx->last.bin_total;
            keep.heavy_total = ctx->last.heavy_total;
            CUDA_TRY(cudaMemcpyAsync(ctx->status.p, &keep, sizeof(StepStatus), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        InteractParams I;
        I.m = ctx->dm; I.solids = ctx->solids.p; I.shapes = ctx->shapes.p; I.n_solids = n_solids; I.grid = g;
        I.bin_off = ctx->bin_off.p; I.bin_list = ctx->bin_list.p; I.global_list = ctx->global_list.p; I.U = dU;
        I.dtINV = 1.0 / dt; I.rhof = rhof; I.As = dAs; I.Fs = dFs; I.Ts = dTs; I.Ct = dCt; I.force_torque = dFT;
        I.pair_counts = ctx->pair_counts.p; I.slots = ctx->slots.p; I.K = ctx->K;
        I.vols = ctx->vols.p; I.n_item = ctx->n_item.p; I.heavy = ctx->heavy.p;
        I.heavy_count = &ctx->status.p->heavy_total; I.heavy_cap = (long long)ctx->heavy.n; I.n_global = 0;
        I.excluded = replay ? ctx->excluded.p : nullptr; I.status = ctx->status.p;
        CUDA_TRY(cudaEventRecord(ctx->ev[1], st));
        if (!replay) {
            k_classify<<<grid_for(nC, 256), 256, 0, st>>>(I);
            CUDA_TRY(cudaEventRecord(ctx->ev[2], st));
            const int hgrid = ctx->n_sm * 4;
            if (ctx->dm.is_hex) k_heavy<true><<<hgrid, TPB, 0, st>>>(I);
            else k_heavy<false><<<hgrid, TPB, 0, st>>>(I);
            ctx->launches += 2;
        } else {
            CUDA_TRY(cudaEventRecord(ctx->ev[2], st));   // replay re-uses the classified and evaluated slots
        }
        CUDA_TRY(cudaEventRecord(ctx->ev[3], st));
        k_accumulate<<<grid_for(nC, 256), 256, 0, st>>>(I);
        CUDA_TRY(cudaEventRecord(ctx->ev[4], st));
        ++ctx->launches;
        if (replay) { k_fix_ct<<<grid_for(nC, 256), 256, 0, st>>>(dCt, nC); ++ctx->launches; }
        else {
            ConnParams C;
            C.m = ctx->dm; C.solids = ctx->solids.p; C.Ct = dCt; C.slots = ctx->slots.p; C.K = ctx->K; C.root_count = ctx->root_count.p;
            k_connectivity<<<grid_for(nC, 256), 256, 0, st>>>(C);
            ++ctx->launches;
        }
        k_finalize<<<std::min(grid_for(n_solids, 256), 296), 256, 0, st>>>(ctx->pair_counts.p, ctx->root_count.p, n_solids, ctx->status.p);
        k_scale_ft<<<grid_for(6LL * n_solids, 256), 256, 0, st>>>(dFT, 6 * n_solids, rhof);
        ctx->launches += 2;
        CUDA_TRY(cudaEventRecord(ctx->ev[5], st));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(ctx->h_status, ctx->status.p, sizeof(StepStatus), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        {
            const double add = replay ? 1.0 : 0.0; // a replay pass adds to the first pass of the same call
            for (int k = 0; k < 5; ++k) {
                float x = 0;
                cudaEventElapsedTime(&x, ctx->ev[k], ctx->ev[k + 1]);
                ctx->t_ms[k] = add * ctx->t_ms[k] + x;
            }
            float d = 0;
            cudaEventElapsedTime(&d, ctx->ev[0], ctx->ev[5]);
            ctx->t_ms[5] = add * ctx->t_ms[5] + d;
        }
        {
            const StepStatus prev = ctx->last;
            ctx->last = *ctx->h_status;
            if (replay) { ctx->last.n_flagged = (int)ctx->flagged_last; ctx->last.heavy_total = prev.heavy_total; }
        }
        if (!replay && ctx->last.heavy_total > (unsigned long long)ctx->heavy.n && attempt == 0) {
            CUDA_TRY(ctx->heavy.ensure((size_t)(ctx->last.heavy_total + ctx->last.heavy_total / 4 + 1024)));
            continue;
        }
        if (ctx->last.bin_overflow && !replay && attempt == 0) {
            CUDA_TRY(ctx->bin_list.ensure((size_t)ctx->last.bin_total + (size_t)ctx->last.bin_total / 4 + 1024));
            continue;
        }
        break;
    }
    if (ctx->last.bin_overflow) return fail(SDFIBM_ERR_CAPACITY, "solid bin list overflow");
    if (ctx->last.bad_cell) return fail(SDFIBM_ERR_UNSUPPORTED, "cell with more than 32 vertices");
    if (ctx->last.slot_overflow)
        return fail(SDFIBM_ERR_CAPACITY, "more solids touch one cell than the slot count; raise it with sdfibm_set_cell_slots");
    return SDFIBM_OK;
}

static int ensure_fields(sdfibm_context *ctx, int n_solids) {
    const size_t nC = ctx->dm.n_cells;
    CUDA_TRY(ctx->dU.ensure(3 * nC));
    CUDA_TRY(ctx->dAs.ensure(nC));
    CUDA_TRY(ctx->dFs.ensure(3 * nC));
    CUDA_TRY(ctx->dTs.ensure(nC));
    CUDA_TRY(ctx->dCt.ensure(nC));
    CUDA_TRY(ctx->dFT.ensure(6 * (size_t)n_solids));
    return SDFIBM_OK;
}

int sdfibm_interact(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, const double *U, double dt, double rhof,
                    double *As, double *Fs, double *Ts, double *Ct, double *force_torque) {
    if (!ctx || !solids || n_solids <= 0 || !U || !As || !Fs || !Ts || !Ct || !force_torque)
        return fail(SDFIBM_ERR_ARG, "sdfibm_interact: null argument");
    if (!ctx->has_mesh || ctx->h_shapes.empty()) return fail(SDFIBM_ERR_STATE, "sdfibm_interact: set mesh and shapes first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_fields(ctx, n_solids);
    if (rc) return rc;
    const size_t nC = ctx->dm.n_cells;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(ctx->dU.p, U, sizeof(double) * 3 * nC, cudaMemcpyHostToDevice, st));
    rc = sdfibm_interact_device(ctx, solids, n_solids, ctx->dU.p, dt, rhof, ctx->dAs.p, ctx->dFs.p, ctx->dTs.p, ctx->dCt.p, ctx->dFT.p);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(As, ctx->dAs.p, sizeof(double) * nC, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(Fs, ctx->dFs.p, sizeof(double) * 3 * nC, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(Ts, ctx->dTs.p, sizeof(double) * nC, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(Ct, ctx->dCt.p, sizeof(double) * nC, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(force_torque, ctx->dFT.p, sizeof(double) * 6 * n_solids, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SDFIBM_OK;
}

int sdfibm_fix_internal_device(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, double *dU, const double *dCt) {
    if (!ctx || !solids || n_solids <= 0 || !dU) return fail(SDFIBM_ERR_ARG, "sdfibm_fix_internal: null argument");
    if (!ctx->has_mesh) return fail(SDFIBM_ERR_STATE, "sdfibm_fix_internal: set mesh first");
    if (!dCt) dCt = ctx->last_Ct;
    if (!dCt) return fail(SDFIBM_ERR_STATE, "sdfibm_fix_internal: no Ct available (call interact first)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = stage_solids(ctx, solids, n_solids);
    if (rc) return rc;
    k_fix_internal<<<grid_for(ctx->dm.n_cells, 256), 256, 0, ctx->stream>>>(ctx->dm, ctx->solids_in.p, n_solids, dCt, dU);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SDFIBM_OK;
}

int sdfibm_fix_internal(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, double *U) {
    if (!ctx || !U) return fail(SDFIBM_ERR_ARG, "sdfibm_fix_internal: null argument");
    if (!ctx->has_mesh || !ctx->dCt.p || ctx->last_Ct != ctx->dCt.p)
        return fail(SDFIBM_ERR_STATE, "sdfibm_fix_internal: call sdfibm_interact (host buffers) first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t nC = ctx->dm.n_cells;
    CUDA_TRY(cudaMemcpyAsync(ctx->dU.p, U, sizeof(double) * 3 * nC, cudaMemcpyHostToDevice, ctx->stream));
    int rc = sdfibm_fix_internal_device(ctx, solids, n_solids, ctx->dU.p, ctx->dCt.p);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(U, ctx->dU.p, sizeof(double) * 3 * nC, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SDFIBM_OK;
}

int sdfibm_candidate_counts(sdfibm_context *ctx, int64_t counts[3]) {
    if (!ctx || !counts) return fail(SDFIBM_ERR_ARG, "null argument");
    for (int k = 0; k < 3; ++k) counts[k] = (int64_t)ctx->last.counts[k];
    return SDFIBM_OK;
}

int sdfibm_last_stats(sdfibm_context *ctx, int64_t stats[4]) {
    if (!ctx || !stats) return fail(SDFIBM_ERR_ARG, "null argument");
    stats[0] = ctx->flagged_last;
    stats[1] = ctx->launches;
    stats[2] = ctx->last.bin_total;
    stats[3] = (int64_t)ctx->last.heavy_total;
    return SDFIBM_OK;
}

int sdfibm_last_timings(sdfibm_context *ctx, double ms[6]) {
    if (!ctx || !ms) return fail(SDFIBM_ERR_ARG, "null argument");
    for (int k = 0; k < 6; ++k) ms[k] = ctx->t_ms[k];
    return SDFIBM_OK;
}

int sdfibm_candidate_lists(sdfibm_context *ctx, int32_t *offsets, int32_t *cells, int64_t capacity) {
    if (!ctx || !offsets) return fail(SDFIBM_ERR_ARG, "null argument");
    if (!ctx->last_Ct) return fail(SDFIBM_ERR_STATE, "no interact has run on this context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int nC = ctx->dm.n_cells, nS = ctx->n_solids_last;
    const unsigned char *excl = ctx->last_used_replay ? ctx->excluded.p : nullptr;
    // per (solid,type) sizes come from the pair counters of the last pass
    std::vector<unsigned> h_cnt(3 * (size_t)nS);
    CUDA_TRY(cudaMemcpyAsync(h_cnt.data(), ctx->pair_counts.p, sizeof(unsigned) * 3 * nS, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int64_t total = 0;
    offsets[0] = 0;
    for (size_t k = 0; k < h_cnt.size(); ++k) { total += h_cnt[k]; offsets[k + 1] = (int32_t)total; }
    if (!cells) return SDFIBM_OK;
    if (total > capacity) return fail(SDFIBM_ERR_CAPACITY, "candidate list capacity too small");
    if (total == 0) return SDFIBM_OK;
    DevBuf<int> cnt, off, vals, vals2;
    DevBuf<unsigned> keys, keys2;
    DevBuf<unsigned char> tmp;
    CUDA_TRY(cnt.ensure((size_t)nC + 1));
    CUDA_TRY(off.ensure((size_t)nC + 1));
    CUDA_TRY(keys.ensure(total)); CUDA_TRY(keys2.ensure(total)); CUDA_TRY(vals.ensure(total)); CUDA_TRY(vals2.ensure(total));
    CUDA_TRY(cudaMemsetAsync(cnt.p, 0, sizeof(int) * ((size_t)nC + 1), st));
    k_list_count<<<grid_for(nC, 256), 256, 0, st>>>(ctx->last_Ct, ctx->slots.p, excl, ctx->K, nC, cnt.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, off.p, nC + 1, st);
    CUDA_TRY(tmp.ensure(tb));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, off.p, nC + 1, st);
    k_list_emit<<<grid_for(nC, 256), 256, 0, st>>>(ctx->last_Ct, ctx->slots.p, excl, ctx->K, nC, off.p, keys.p, vals.p);
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, keys.p, keys2.p, vals.p, vals2.p, (int)total, 0, 32, st);
    CUDA_TRY(tmp.ensure(sb));
    cub::DeviceRadixSort::SortPairs(tmp.p, sb, keys.p, keys2.p, vals.p, vals2.p, (int)total, 0, 32, st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(cells, vals2.p, sizeof(int) * total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    cnt.release(); off.release(); vals.release(); vals2.release(); keys.release(); keys2.release(); tmp.release();
    return SDFIBM_OK;
}

int sdfibm_collide(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n, double delta, int32_t *pairs, int64_t cap,
                   int64_t *n_pairs, double *force_torque) {
    if (!ctx || !solids || n <= 0 || !n_pairs) return fail(SDFIBM_ERR_ARG, "sdfibm_collide: null argument");
    if (!ctx->has_mesh || ctx->h_shapes.empty()) return fail(SDFIBM_ERR_STATE, "sdfibm_collide: set mesh and shapes first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    *n_pairs = 0;
    UGridDev g;                                                                // UGrid ctor, ugrid.cpp:5-22
    g.deltaINV = 1.0 / delta;
    for (int d = 0; d < 3; ++d) g.lo[d] = ctx->bmin[d];
    g.nx = (int)std::ceil((ctx->bmax[0] - ctx->bmin[0]) * g.deltaINV);
    g.ny = (int)std::ceil((ctx->bmax[1] - ctx->bmin[1]) * g.deltaINV);
    g.nz = (int)std::ceil((ctx->bmax[2] - ctx->bmin[2]) * g.deltaINV);
    g.nynz = g.ny * g.nz;
    if (g.nx <= 0 || g.ny <= 0 || g.nz <= 0) return SDFIBM_OK;                 // HEAD: delta = -2 -> no pairs
    int rc = stage_solids(ctx, solids, n);
    if (rc) return rc;
    DevBuf<int> ids, sids, pcnt, poff, dpairs;
    DevBuf<unsigned> keys, skeys;
    DevBuf<unsigned char> tmp;
    DevBuf<double> dft;
    CUDA_TRY(keys.ensure(n)); CUDA_TRY(ids.ensure(n)); CUDA_TRY(skeys.ensure(n)); CUDA_TRY(sids.ensure(n));
    CUDA_TRY(pcnt.ensure((size_t)n + 1)); CUDA_TRY(poff.ensure((size_t)n + 1));
    k_col_keys<<<grid_for(n, 256), 256, 0, st>>>(ctx->solids_in.p, n, g, keys.p, ids.p);
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, keys.p, skeys.p, ids.p, sids.p, n, 0, 32, st);
    CUDA_TRY(tmp.ensure(sb));
    cub::DeviceRadixSort::SortPairs(tmp.p, sb, keys.p, skeys.p, ids.p, sids.p, n, 0, 32, st); // stable; keys are signed but compared as bits:
    CUDA_TRY(cudaMemsetAsync(pcnt.p, 0, sizeof(int) * ((size_t)n + 1), st));
    k_col_pairs<<<grid_for(n, 128), 128, 0, st>>>(skeys.p, sids.p, n, g, pcnt.p, nullptr, nullptr, 0, 0);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, pcnt.p, poff.p, n + 1, st);
    CUDA_TRY(tmp.ensure(tb));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, pcnt.p, poff.p, n + 1, st);
    int total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, poff.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *n_pairs = total;
    if (total > 0) {
        CUDA_TRY(dpairs.ensure(2 * (size_t)total));
        k_col_pairs<<<grid_for(n, 128), 128, 0, st>>>(skeys.p, sids.p, n, g, pcnt.p, poff.p, dpairs.p, total, 1);
        if (force_torque) {
            CUDA_TRY(dft.ensure(6 * (size_t)n));
            CUDA_TRY(cudaMemcpyAsync(dft.p, force_torque, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, st));
            k_col_narrow<<<grid_for(total, 128), 128, 0, st>>>(ctx->solids_in.p, ctx->shapes.p, dpairs.p, total, dft.p);
            CUDA_TRY(cudaMemcpyAsync(force_torque, dft.p, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, st));
        }
        if (pairs) {
            if (total > cap) { cudaStreamSynchronize(st); return fail(SDFIBM_ERR_CAPACITY, "collision pair capacity too small"); }
            CUDA_TRY(cudaMemcpyAsync(pairs, dpairs.p, sizeof(int) * 2 * (size_t)total, cudaMemcpyDeviceToHost, st));
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    keys.release(); ids.release(); skeys.release(); sids.release(); pcnt.release(); poff.release(); dpairs.release(); tmp.release(); dft.release();
    return SDFIBM_OK;
}

} // extern "C"
