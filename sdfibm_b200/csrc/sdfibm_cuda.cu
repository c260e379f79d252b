// sdfibm_cuda.cu — sm_100a kernels and the C-ABI implementation of the coupling path.
//
// Design: the reference walks solids and flood-fills cells per solid
// (src/solidcloud.cpp:445-446, src/cellenumerator.cpp:6-34).  Here the loop is inverted: ONE thread
// per mesh cell walks the few solids whose (conservatively inflated) bounding volume covers that
// cell, taken from a per-step uniform-grid binning of the solids.  Every field (As, Fs, Ts, Ct) is
// then written exactly once per cell, fully coalesced, with the per-cell sums taken in ascending
// solid order exactly like the reference's `+=` over its solid loop — no field memsets, no field
// atomics.  The flood fill's "face-connected component of the seed" semantics is
// restored by a connectivity certificate fused into the kernels and, only for solids that fail it, an
// exact check plus label-propagation replay (all on the GPU).
//
// Compiled with -fmad=false: predicates are strict `<` on un-contracted fp64.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is bound at run time (sdfibm_comm_*), a host that never calls them needs no NCCL
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
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
#define HEX_NONE 0xffffffffu

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
    float pos32[3]; // centre relative to the mesh origin (keys of the connectivity certificate)
    int conn_proven; // the vertex-inside cell set of this solid is provably face connected: no certificate needed
    int ball_fast;   // an un-rotated sphere: k_heavy_box's separable corner path applies (the queue keeps such pairs apart from the rest)
    float ri32;     // KIND_3D: certified inner radius minus the fp32 slack, rounded down (<= 0: none): the same test as k_classify
    // fp32 refinement of the pre-classification for convex analytic shapes (k_classify): body coordinates b = M (p - pos) + com;
    // refine 1: inside <=> sum (b_i rp_i)^2 < 1 (ellipsoid, ellipse: rp = 1/radius, 0 for an ignored axis);
    // refine 2: inside <=> |b_i| < rp_i for every i (box, rectangle: rp = half width, huge for an ignored axis); 0: none
    float M[9], rp[3], com32[3];
    float eps_ref;  // certified bound of the fp32 error of the refinement function near the surface
    int refine;
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
    const double2 *face_rec; // hex path: 64-byte record per face: Cf.xyz, Sf.xyz, |Sf|, pad
    const unsigned *hex_topo; // hex meshes: 3 words per cell, 4-bit cell-local vertex slot of every face vertex
    const int *nb6;           // per cell: for each signed axis direction (-x,+x,-y,+y,-z,+z) the best-aligned face neighbour (-1 = none)
    // Cells are renumbered at upload into TILE ORDER (position i = rank of the cell in a stable sort by the key of the
    // spatial tile holding its centre; ~256 cells per tile).  Every per-cell array above/below is stored in position
    // order; orig[i] is the caller's cell label of position i (U is read and As/Fs/Ts/Ct are written through it).
    const int *orig;          // position -> caller's cell label
    const float4 *cc32;       // per position: fp32 (centre - origin, vertex-cloud radius rounded up): conservative pre-classification
    const unsigned *tile_key; // per position: linear index of its tile in the tile grid (= solid-bin index)
    double origin[3];         // centre of the mesh bounds
    // Bounding box of each cell's vertex cloud about its centre (half extents, rounded up; w = 1 when the cell IS that box: every
    // vertex sits on a corner).  k_classify bounds the nearest / farthest vertex of a cell with it: for box cells the
    // pre-classification is then exact up to the fp32 slack, so only cells that really have vertices on both sides are queued.
    const float4 *cell_box;   // per position (nullptr when box_uniform)
    float4 box_const;         // the one box of a uniform mesh
    int box_uniform;
    int V_uniform;            // every cell has bit for bit the same volume: k_final takes V_const instead of 8 bytes per touched cell
    double V_const;
    // the mesh is a complete nx x ny x nz lattice of identical boxes whose lattice neighbours are all face neighbours: for such a
    // mesh the vertex-inside cell set of a ball that lies inside the mesh is provably face connected (see k_solid_prepare)
    int lattice_full;
    double lattice_lo[3], lattice_hi[3];
    float2 rad_const;       // upper bound of cell_rad over the mesh (used for every cell when the mesh is near uniform)
    int rad_uniform;
    int is_hex;             // every cell has 8 points / 6 faces and every face 4 points
    int mixed;              // some cells are such hexahedra and some are not: hex_topo[3c] == HEX_NONE marks the others, which go
                            // through the general-polyhedron kernel from a queue of their own
    int two_d;
    // Exact-box meshes (checked cell by cell at upload, k_box_topo): every cell's 8 vertices sit bit for bit on the corners of
    // an axis-aligned box, and every face record agrees bit for bit with its vertices (Cf on the face plane, Sf along its
    // normal).  The exact evaluation then needs 6 coordinates and two topology words per cell instead of 8 gathered vertices
    // and 6 face records, and its face areas need no cross products or square roots (k_heavy_box).
    int box_exact;            // 0: no; 1: yes, face-plane coordinates from cfa6; 2: yes, and every Cf lies bit for bit on its plane
    const double *box6;       // per position: lo.xyz, hi.xyz of the cell
    const double *cfa6;       // per position (box_exact == 1): normal coordinate of the mesh's Cf of face (axis a, side s) at [2a+s]
    const unsigned *btopo;    // per position: [0] corner code (bit0 x-hi, bit1 y-hi, bit2 z-hi) of cellPoints() slot m at bits 3m;
                              // [1] for face (axis a, side s) at bits 3(2a+s): in-plane corner (iu, iv) of the face loop's
                              // first vertex + loop direction (0: second vertex differs along u = (a+1)%3, 1: along v = (a+2)%3)
};

// the tile grid: static per mesh; solids are binned on it every step
struct BinGrid {
    double lo[3];
    double inv[3];
    int n[3];
    int n_bins;
};

struct StepStatus {
    unsigned long long counts[3]; // ALL_INSIDE, CENTER_INSIDE, CENTER_OUTSIDE pairs
    int n_flagged;                // solids with >1 roots in the exact connectivity check
    int slot_overflow;            // a cell was touched by more than K solids
    int bin_overflow;             // bin list capacity exceeded
    int bad_cell;                 // a cell/face exceeded MAX_CELL_VERTS / MAX_FACE_VERTS
    int bad_shape;                // a solid refers to a shape index outside the table
    int bin_total;
    int n_global;
    unsigned long long heavy_total; // (cell, solid) items that needed exact evaluation (mixed meshes: those of hexahedral cells)
    unsigned long long heavy_gen;   // mixed meshes: items of non-hexahedral cells, queued from the back of the same buffer
    int need_cert;                  // some binned solid's cell set is not connected by construction: the certificate pass has work
    int slot_need;                  // slot_overflow: the largest number of solids that touch one cell
};

__device__ __forceinline__ D3 ld3(const double *__restrict__ p, long long i) {
    return {__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)};
}

__device__ __forceinline__ int bin_coord(const BinGrid &g, double x, int d) {
    // monotone in x (the same expression maps cells and solid boxes); __double2int_rd saturates, so +-1e300 is safe
    const int i = __double2int_rd((x - g.lo[d]) * g.inv[d]);
    return min(max(i, 0), g.n[d] - 1);
}

// ------------------------------------------------------------------------------------------------
// K0  per-cell vertex-cloud radii (once per mesh)
// ------------------------------------------------------------------------------------------------
__global__ void k_face_mag(const double *Cf, const double *Sf, int n_faces, double *magSf, double2 *face_rec) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const D3 S = ld3(Sf, f);
    const double mg = mag3(S);
    magSf[f] = mg;
    if (face_rec) {
        const D3 C = ld3(Cf, f);
        face_rec[4 * (long long)f] = make_double2(C.x, C.y);
        face_rec[4 * (long long)f + 1] = make_double2(C.z, S.x);
        face_rec[4 * (long long)f + 2] = make_double2(S.y, S.z);
        face_rec[4 * (long long)f + 3] = make_double2(mg, 0.0);
    }
}

// hex meshes: for face k (cells() order) of cell c and its j-th vertex (faces() order), the position of that
// vertex in the cell's cellPoints() list, packed 4 bits each: word k/2, bit 16*(k&1) + 4*j.
__global__ void k_hex_topo(DevMesh m, unsigned *topo, int *bad) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.n_cells) return;
    const int pb = m.cp_off[c], fb = m.cf_off[c];
    bool hex = (m.cp_off[c + 1] - pb == 8) && (m.cf_off[c + 1] - fb == 6);
    for (int k = 0; k < 6 && hex; ++k) { const int f = m.cf[fb + k]; hex = m.fp_off[f + 1] - m.fp_off[f] == 4; }
    if (!hex) {   // not a hexahedron with quadrilateral faces: general path
        topo[3 * (long long)c] = HEX_NONE; topo[3 * (long long)c + 1] = 0u; topo[3 * (long long)c + 2] = 0u;
        return;
    }
    int vid[8];
    for (int k = 0; k < 8; ++k) vid[k] = m.cp[pb + k];
    unsigned w[3] = {0u, 0u, 0u};
    for (int k = 0; k < 6; ++k) {
        const int f = m.cf[fb + k];
        for (int j = 0; j < 4; ++j) {
            const int g = m.fp[m.fp_off[f] + j];
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


// Exact-box topology of a cell (see DevMesh::box_exact).  Any deviation clears *ok for the whole mesh.
__global__ void k_box_topo(DevMesh m, double *box6, unsigned *btopo, double *cfa6, int *ok) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.n_cells) return;
    const int pb = m.cp_off[c], fb = m.cf_off[c];
    if (m.cp_off[c + 1] - pb != 8 || m.cf_off[c + 1] - fb != 6) { atomicExch(ok, 0); return; }
    int vid[8];
    D3 p[8];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < 8; ++k) {
        vid[k] = m.cp[pb + k];
        p[k] = ld3(m.points, vid[k]);
        lo[0] = fmin(lo[0], p[k].x); lo[1] = fmin(lo[1], p[k].y); lo[2] = fmin(lo[2], p[k].z);
        hi[0] = fmax(hi[0], p[k].x); hi[1] = fmax(hi[1], p[k].y); hi[2] = fmax(hi[2], p[k].z);
    }
    bool good = lo[0] < hi[0] && lo[1] < hi[1] && lo[2] < hi[2];
    int code[8];
    unsigned seen = 0, w0 = 0, w1 = 0;
    for (int k = 0; k < 8; ++k) {
        const double q[3] = {p[k].x, p[k].y, p[k].z};
        int cd = 0;
        for (int a = 0; a < 3; ++a) {
            if (q[a] == hi[a]) cd |= 1 << a;
            else if (q[a] != lo[a]) good = false;
        }
        code[k] = cd;
        seen |= 1u << cd;
        w0 |= (unsigned)cd << (3 * k);
    }
    good = good && seen == 0xffu;
    unsigned fseen = 0;
    bool cf_exact = true;
    double cfa[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6 && good; ++k) {
        const int f = m.cf[fb + k];
        const int q0 = m.fp_off[f];
        if (m.fp_off[f + 1] - q0 != 4) { good = false; break; }
        int fc[4];
        for (int j = 0; j < 4; ++j) {
            const int g = m.fp[q0 + j];
            int l = -1;
            for (int t = 0; t < 8; ++t) if (vid[t] == g) l = t;
            if (l < 0) { good = false; l = 0; }
            fc[j] = code[l];
        }
        const int andc = fc[0] & fc[1] & fc[2] & fc[3], orc = fc[0] | fc[1] | fc[2] | fc[3];
        const int cm = ~(andc ^ orc) & 7;   // the axes along which the four corners agree
        if (__popc(cm) != 1) { good = false; break; }
        const int a = __ffs(cm) - 1, s = (fc[0] >> a) & 1, u = (a + 1) % 3, v = (a + 2) % 3;
        const int d1 = fc[0] ^ fc[1];
        int dir;
        if (d1 == (1 << u)) dir = 0;
        else if (d1 == (1 << v)) dir = 1;
        else { good = false; break; }
        if (fc[2] != (fc[0] ^ (1 << u) ^ (1 << v)) || fc[3] != (fc[0] ^ (dir ? (1 << u) : (1 << v)))) { good = false; break; }
        const int fi = 2 * a + s;
        fseen |= 1u << fi;
        w1 |= (unsigned)(((fc[0] >> u) & 1) | (((fc[0] >> v) & 1) << 1) | (dir << 2)) << (3 * fi);
        // the face record agrees bit for bit with the vertices: centre on the face plane, area vector along the normal
        const double Cf[3] = {m.Cf[3 * (long long)f], m.Cf[3 * (long long)f + 1], m.Cf[3 * (long long)f + 2]};
        const double Sf[3] = {m.Sf[3 * (long long)f], m.Sf[3 * (long long)f + 1], m.Sf[3 * (long long)f + 2]};
        const double A = (hi[u] - lo[u]) * (hi[v] - lo[v]);
        if (Sf[u] != 0.0 || Sf[v] != 0.0 || !(fabs(fabs(Sf[a]) - A) <= 1e-14 * A)) good = false;
        // ... and the centre lies within rounding of the face plane; where it is not bit for bit ON the plane the kernel takes
        // the plane coordinate of the pyramid height from the mesh's own Cf (cfa6) instead of from the vertices
        const double plane = s ? hi[a] : lo[a];
        if (!(fabs(Cf[a] - plane) <= 1e-12 * (hi[a] - lo[a]))) good = false;
        if (Cf[a] != plane) cf_exact = false;
        cfa[fi] = Cf[a];
    }
    good = good && fseen == 0x3fu;
    if (!good) { atomicExch(ok, 0); return; }
    if (!cf_exact) atomicExch(ok + 1, 0);
    for (int k = 0; k < 6; ++k) cfa6[6 * (long long)c + k] = cfa[k];
    for (int a = 0; a < 3; ++a) { box6[6 * (long long)c + a] = lo[a]; box6[6 * (long long)c + 3 + a] = hi[a]; }
    btopo[2 * (long long)c] = w0;
    btopo[2 * (long long)c + 1] = w1;
}

// direction table for the connectivity certificate: which face neighbour lies towards -x, +x, -y, ...
__global__ void k_nb6(DevMesh m, int *nb6) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.n_cells) return;
    const D3 cc = ld3(m.cc, c);
    const int b = m.nb_off[c], e = m.nb_off[c + 1];
    for (int d = 0; d < 6; ++d) {
        int best = -1;
        double best_cos = 0.0;
        for (int k = b; k < e; ++k) {
            const D3 r = ld3(m.cc, m.nb[k]) - cc;
            const double comp = (d >> 1) == 0 ? r.x : (d >> 1) == 1 ? r.y : r.z;
            const double cs = ((d & 1) ? comp : -comp) / (mag3(r) + 1e-300);
            if (cs > best_cos) { best_cos = cs; best = m.nb[k]; }
        }
        nb6[6 * (long long)c + d] = best;
    }
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

// per-cell half extents of the vertex cloud + "is an axis-aligned box" flag; ext[0..2] = max, ext[3..5] = min over the mesh
// (as int bits of non-negative floats), ext[6] = number of non-box cells
__global__ void k_v_uniform(const double *V, int n_cells, int *differs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cells && __double_as_longlong(V[c]) != __double_as_longlong(V[0])) *differs = 1;
}

__global__ void k_cell_box(DevMesh m, float4 *box, int *ext) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.n_cells) return;
    const D3 cc = ld3(m.cc, c);
    const int b = m.cp_off[c], e = m.cp_off[c + 1];
    double hx = 0.0, hy = 0.0, hz = 0.0;
    for (int k = b; k < e; ++k) {
        const D3 d = ld3(m.points, m.cp[k]) - cc;
        hx = fmax(hx, fabs(d.x)); hy = fmax(hy, fabs(d.y)); hz = fmax(hz, fabs(d.z));
    }
    bool is_box = (e - b == 8);
    for (int k = b; k < e && is_box; ++k) {
        const D3 d = ld3(m.points, m.cp[k]) - cc;
        is_box = fabs(d.x) >= hx * (1.0 - 1e-9) && fabs(d.y) >= hy * (1.0 - 1e-9) && fabs(d.z) >= hz * (1.0 - 1e-9);
    }
    const float fx = __double2float_ru(hx * (1.0 + REL_MARGIN)), fy = __double2float_ru(hy * (1.0 + REL_MARGIN)), fz = __double2float_ru(hz * (1.0 + REL_MARGIN));
    box[c] = make_float4(fx, fy, fz, is_box ? 1.f : 0.f);
    atomicMax(ext + 0, __float_as_int(fx)); atomicMax(ext + 1, __float_as_int(fy)); atomicMax(ext + 2, __float_as_int(fz));
    atomicMin(ext + 3, __float_as_int(fx)); atomicMin(ext + 4, __float_as_int(fy)); atomicMin(ext + 5, __float_as_int(fz));
    if (!is_box) atomicAdd(ext + 6, 1);
}

// ------------------------------------------------------------------------------------------------
// K0b  tile-order renumbering of the cells (once per mesh)
// ------------------------------------------------------------------------------------------------
__global__ void k_tile_keys(const double *cc, int n_cells, BinGrid g, unsigned *keys, int *ids) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const D3 x = ld3(cc, c);
    keys[c] = (unsigned)((bin_coord(g, x.z, 2) * g.n[1] + bin_coord(g, x.y, 1)) * g.n[0] + bin_coord(g, x.x, 0));
    ids[c] = c;
}
__global__ void k_perm_inverse(const int *orig, int n_cells, int *inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_cells) inv[orig[i]] = i;
}
__global__ void k_perm_counts(const int *off, const int *orig, int n_cells, int *cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_cells) return;
    cnt[i] = (i < n_cells) ? off[orig[i] + 1] - off[orig[i]] : 0;
}
// vals_p[off_p[i] + k] = map(vals[off[orig[i]] + k])
__global__ void k_perm_csr(const int *off, const int *vals, const int *orig, const int *off_p, const int *map, int n_cells, int *vals_p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const int b = off[orig[i]], n = off[orig[i] + 1] - b, o = off_p[i];
    for (int k = 0; k < n; ++k) {
        const int v = vals[b + k];
        vals_p[o + k] = map ? map[v] : v;
    }
}
__global__ void k_perm_cells(const double *cc, const double *V, const int *orig, int n_cells, double *cc_p, double *V_p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const int c = orig[i];
    cc_p[3 * (long long)i] = cc[3 * (long long)c];
    cc_p[3 * (long long)i + 1] = cc[3 * (long long)c + 1];
    cc_p[3 * (long long)i + 2] = cc[3 * (long long)c + 2];
    V_p[i] = V[c];
}
// fp32 copy of the centres relative to the mesh origin + the per-cell vertex-cloud radius (already rounded up)
__global__ void k_cc32(const double *cc_p, const float2 *rad, int n_cells, double ox, double oy, double oz, float4 *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    out[i] = make_float4((float)(cc_p[3 * (long long)i] - ox), (float)(cc_p[3 * (long long)i + 1] - oy),
                         (float)(cc_p[3 * (long long)i + 2] - oz), rad[i].x);
}
// [cmin, cmax] of the caller's cell labels over each chunk of positions (host-buffer pipeline)
__global__ void k_chunk_ranges(const int *orig, int n_cells, int n_chunk, int *cmin, int *cmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // chunk k covers positions [n*k/n_chunk, n*(k+1)/n_chunk)
    int k = -1, v = 0;
    if (i < n_cells) {
        k = (int)(((long long)i * n_chunk) / n_cells);
        while ((long long)n_cells * k / n_chunk > i) --k;
        while ((long long)n_cells * (k + 1) / n_chunk <= i) ++k;
        v = orig[i];
    }
    // a warp lies inside one chunk except at the few chunk boundaries: reduce in the warp, one atomic pair per warp
    const unsigned FULL = 0xffffffffu;
    const int k0 = __shfl_sync(FULL, k, 0);
    if (__all_sync(FULL, k == k0) && k0 >= 0) {
        const int lo = __reduce_min_sync(FULL, v), hi = __reduce_max_sync(FULL, v);
        if ((threadIdx.x & 31) == 0) { atomicMin(cmin + k0, lo); atomicMax(cmax + k0, hi); }
    } else if (k >= 0) {
        atomicMin(cmin + k, v);
        atomicMax(cmax + k, v);
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
    double origin[3];
    double half_ext, rad_max;   // fp32 slack of the conservative tests: 4e-6 (half_ext + r_out + rad_max)
    int lattice_full;
    double box_h[3];            // half extents of the lattice cell
    int *bin_count;   // [n_bins+1]
    int *global_list; // [n_solids]
    StepStatus *status;
    int fixed_cap;    // > 0: every tile owns fixed_cap list slots (no scan, no separate fill pass): the solid is appended right here
    int *bin_list;    // [n_bins * fixed_cap] in that mode
};

// SUB threads per solid share the loop over the bins its bounding box covers
#define BIN_SUB 8
__device__ __forceinline__ int sh_kind_of(const DevShape &sh) { return sh.kind; }
__global__ void k_solid_prepare(PrepParams P) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = gid / BIN_SUB, sub = gid % BIN_SUB;
    if (s >= P.n_solids) return;
    sdfibm_solid_t in = P.solids[s];
    if (in.shape < 0 || in.shape >= P.n_shapes) { P.status->bad_shape = 1; in.shape = 0; }
    DevSolid S;
    for (int d = 0; d < 3; ++d) { S.pos[d] = in.pos[d]; S.vel[d] = in.vel[d]; S.omega[d] = in.omega[d]; }
    for (int d = 0; d < 4; ++d) S.q[d] = in.quat[d];
    for (int d = 0; d < 3; ++d) S.pos32[d] = (float)(in.pos[d] - P.origin[d]);
    S.ri32 = (sh_kind_of(P.shapes[in.shape]) == KIND_3D) ? __double2float_rd(P.shapes[in.shape].r_in - 4e-6 * (P.half_ext + P.shapes[in.shape].r_out + P.rad_max)) : 0.f;
    S.shape = in.shape;
    const DevShape &sh = P.shapes[in.shape];
    {
        const sdfibm_shape_t &sp = sh.s;
        const double db = 4e-6 * (P.half_ext + sh.r_out + P.rad_max);   // bound of the fp32 error of a body coordinate (5x margin)
        S.refine = 0;
        S.eps_ref = 0.f;
        for (int d = 0; d < 3; ++d) { S.rp[d] = 0.f; S.com32[d] = (float)sp.com[d]; }
        double rmin = 0.0;
        if (sp.tag == SDFIBM_SHAPE_ELLIPSOID) { S.refine = 1; for (int d = 0; d < 3; ++d) { S.rp[d] = (float)(1.0 / sp.p[d]); S.com32[d] = 0.f; } rmin = fmin(sp.p[0], fmin(sp.p[1], sp.p[2])); }
        else if (sp.tag == SDFIBM_SHAPE_ELLIPSE) { S.refine = 1; S.rp[0] = (float)(1.0 / sp.p[0]); S.rp[1] = (float)(1.0 / sp.p[1]); rmin = fmin(sp.p[0], sp.p[1]); }
        else if (sp.tag == SDFIBM_SHAPE_BOX) { S.refine = 2; for (int d = 0; d < 3; ++d) S.rp[d] = (float)sp.p[d]; }
        else if (sp.tag == SDFIBM_SHAPE_RECTANGLE) { S.refine = 2; S.rp[0] = (float)sp.p[0]; S.rp[1] = (float)sp.p[1]; S.rp[2] = 3.0e38f; }
        if (S.refine == 1) S.eps_ref = __double2float_ru(6.0 * db / rmin + 1e-5);
        if (S.refine == 2) S.eps_ref = __double2float_ru(2.0 * db + 1e-6 * sh.r_out);
        // rows of M: world2local of the unit vectors (the same quaternion sandwich the exact path uses, rounded to fp32)
        const DQ qq = {in.quat[0], {in.quat[1], in.quat[2], in.quat[3]}};
        const D3 zero = {0.0, 0.0, 0.0};
        const D3 cx = world2local(qq, zero, D3{1.0, 0.0, 0.0}), cy = world2local(qq, zero, D3{0.0, 1.0, 0.0}), cz = world2local(qq, zero, D3{0.0, 0.0, 1.0});
        S.M[0] = (float)cx.x; S.M[1] = (float)cy.x; S.M[2] = (float)cz.x;
        S.M[3] = (float)cx.y; S.M[4] = (float)cy.y; S.M[5] = (float)cz.y;
        S.M[6] = (float)cx.z; S.M[7] = (float)cy.z; S.M[8] = (float)cz.z;
        if (!(fabs(magSqr3(cx) - 1.0) < 1e-9)) S.refine = 0;   // a non-unit quaternion scales the body frame: leave it to the exact path
    }
    S.kind = sh.kind;
    S.r_out = sh.r_out;
    S.r_in = sh.r_in;
    DQ q = {in.quat[0], {in.quat[1], in.quat[2], in.quat[3]}};
    D3 ax = qtransform(q, D3{0.0, 0.0, 1.0});
    S.axis[0] = ax.x; S.axis[1] = ax.y; S.axis[2] = ax.z;
    S.axis_is_z = (fabs(ax.x) <= 1e-12 && fabs(ax.y) <= 1e-12) ? 1 : 0;
    S.global = (S.kind == KIND_PLANE || (S.kind == KIND_2D && !S.axis_is_z)) ? 1 : 0;
    {
        // A ball (disc) on a complete lattice of boxes — any radius, clipped by the mesh or not, centre inside or outside it.  Let c* be
        // the point of the mesh box nearest the centre c.  From an inside vertex v, along any axis with |v_i - c*_i| > h_i, the lattice
        // neighbour towards c*_i stays in the mesh and is strictly closer to c (c*_i lies between v_i and c_i), hence inside; the walk
        // ends among the <= 3x3x3 vertices around c*.  Two inside vertices u, w of that block both reach, by changing one coordinate
        // at a time to whichever of u_i, w_i is nearer c_i, the same vertex without ever increasing a |d_i| (the computed |p - c|^2 is
        // monotone in every |d_i|: rounding is monotone).  So the inside vertices are lattice connected, and the member cells — the
        // union of the (clipped) 2x2x2 stars of the inside vertices, stars of lattice neighbours sharing a cell — are face connected:
        // the flood fill returns all of them from any seed and the certificate pass can skip this solid
        // (tests/test_connectivity_proof_cpu.py holds the claim to the oracle's real flood fill).
        const sdfibm_shape_t &sp = sh.s;
        const bool ball = sp.tag == SDFIBM_SHAPE_SPHERE, disc = sp.tag == SDFIBM_SHAPE_CIRCLE && S.axis_is_z;
        bool ok = P.lattice_full && (ball || disc) && sp.com[0] == 0.0 && sp.com[1] == 0.0 && sp.com[2] == 0.0;
        // A well-resolved ellipsoid (ellipse) of ANY orientation, lying inside the lattice.  With f(v) = |A v|^2 (inside: f < 1) and
        // M = A^T A, a lattice step of length h_i along axis i changes f by -+2 h_i (M v)_i + h_i^2 M_ii; some axis has
        // |(M v)_i| >= |M v| / sqrt 3 >= sqrt(lambda_min) sqrt(f) / sqrt 3, so f strictly decreases — by a margin far above rounding —
        // as long as sqrt(f) > rho := (sqrt 3 / 2) h lambda_max / sqrt(lambda_min) = (sqrt 3 / 2) h a_max / a_min^2.  Every inside vertex
        // therefore walks, inside, into the small ellipsoid f <= rho^2; the lattice box around that one lies within f <= 3 rho^2 < 1,
        // so its vertices are all inside and mutually connected.  Required with a 1.5x margin on rho, and the body two cells inside
        // the mesh (the walk must not leave it).
        const bool ellipsoid = sp.tag == SDFIBM_SHAPE_ELLIPSOID, ellipse = sp.tag == SDFIBM_SHAPE_ELLIPSE && S.axis_is_z && sp.com[0] == 0.0 && sp.com[1] == 0.0;
        if (!ok && P.lattice_full && (ellipsoid || ellipse)) {
            const int nd = ellipsoid ? 3 : 2;
            double a_max = 0.0, a_min = 1e300, h = 0.0;
            for (int d = 0; d < nd; ++d) { a_max = fmax(a_max, sp.p[d]); a_min = fmin(a_min, sp.p[d]); h = fmax(h, 2.0 * P.box_h[d]); }
            const double rho = 1.5 * 0.8660254037844386 * h * a_max / (a_min * a_min);
            ok = 3.0 * rho * rho < 0.9;
            for (int d = 0; d < nd; ++d) {
                const double marg = a_max + 4.0 * P.box_h[d] * (1.0 + 1e-5);
                ok = ok && in.pos[d] - marg > P.mesh_lo[d] && in.pos[d] + marg < P.mesh_hi[d];
            }
        }
        S.conn_proven = ok ? 1 : 0;
        S.ball_fast = (sp.tag == SDFIBM_SHAPE_SPHERE && quat_is_identity(q)) ? 1 : 0;
    }
    if (sub == 0) P.out[s] = S;
    if (S.global) {
        if (sub == 0) {
            int k = atomicAdd(&P.status->n_global, 1);
            P.global_list[k] = s;
        }
        return;
    }
    int lo[3], hi[3];
    if (!solid_bin_range(S, P.grid, P.rad3_max, P.radxy_max, P.mesh_lo, P.mesh_hi, lo, hi)) return;
    const int nx = hi[0] - lo[0] + 1, ny = hi[1] - lo[1] + 1, nz = hi[2] - lo[2] + 1;
    for (int t = sub; t < nx * ny * nz; t += BIN_SUB) {
        const int i = lo[0] + t % nx, j = lo[1] + (t / nx) % ny, k = lo[2] + t / (nx * ny);
        const int b = (k * P.grid.n[1] + j) * P.grid.n[0] + i;
        const int pos = atomicAdd(&P.bin_count[b], 1);
        if (P.fixed_cap) {
            if (pos < P.fixed_cap) P.bin_list[(long long)b * P.fixed_cap + pos] = s;
            else P.status->bin_overflow = 2;   // a tile with more candidates than slots: the host falls back to the scan + fill path
        }
    }
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
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = gid / BIN_SUB, sub = gid % BIN_SUB;
    if (s >= P.n_solids) return;
    const DevSolid &S = P.solids[s];
    if (S.global) return;
    int lo[3], hi[3];
    if (!solid_bin_range(S, P.grid, P.rad3_max, P.radxy_max, P.mesh_lo, P.mesh_hi, lo, hi)) return;
    const int nx = hi[0] - lo[0] + 1, ny = hi[1] - lo[1] + 1, nz = hi[2] - lo[2] + 1;
    for (int t = sub; t < nx * ny * nz; t += BIN_SUB) {
        const int i = lo[0] + t % nx, j = lo[1] + (t / nx) % ny, k = lo[2] + t / (nx * ny);
        const int b = (k * P.grid.n[1] + j) * P.grid.n[0] + i;
        const int pos = P.bin_off[b] + atomicAdd(&P.bin_cursor[b], 1);
        if (pos < P.bin_cap) P.bin_list[pos] = s;
        else P.status->bin_overflow = 1;
    }
}

// ascending solid id inside every bin (and the global list): the per-cell accumulation order
__device__ __forceinline__ void bin_insertion_sort(int *lst, int beg, int end) {
    for (int i = beg + 1; i < end; ++i) {
        int v = lst[i], j = i - 1;
        while (j >= beg && lst[j] > v) { lst[j + 1] = lst[j]; --j; }
        lst[j + 1] = v;
    }
}

#ifndef CLS_NT
#define CLS_NT 128   // k_classify (one position per thread): threads per CTA (C5 at 256^3: 0.373 ms at 256, 0.352 at 128, 0.355 at 64)
#endif
#ifndef FINAL_CTAS
#define FINAL_CTAS (1024 / FINAL_NT)   // 1024 resident threads per SM at 64 registers
#endif
#include "interact_kernels.cuh"

// Thread per bin: sort the bin's solids by id, then write the inline fp32 candidate records k_classify reads.  The radii carry
// the fp32 slack: coordinates relative to the mesh origin are bounded by M = half extent + r_out + rad, so the fp32 distance is
// off by < 1e-6 M; slack = 4e-6 M keeps the three-way test conservative (the exact fp64 predicates decide everything it does not).
__device__ __forceinline__ BinEntry make_bin_entry(const DevSolid &S, int s, double ox, double oy, double oz, double half_ext, double rad_max) {
    const double slack = 4e-6 * (half_ext + S.r_out + rad_max);
    BinEntry e;
    e.x = (float)(S.pos[0] - ox); e.y = (float)(S.pos[1] - oy); e.z = (float)(S.pos[2] - oz);
    e.r_out = __double2float_ru(S.r_out + slack);
    e.r_in = __double2float_rd(S.r_in - slack);
    e.s = s;
    e.kind = S.kind;
    e.refine = S.refine | (S.ball_fast << 8);   // bits 0-7: corner-refinement mode; bit 8: un-rotated sphere
    return e;
}

__global__ void k_bin_sort_entries(const int *bin_off, int *bin_list, int n_bins, int bin_cap, int *global_list, StepStatus *status,
                                   const DevSolid *solids, BinEntry *out, double ox, double oy, double oz, double half_ext, double rad_max,
                                   unsigned char *tile_proven, int fixed_cap, const int *bin_count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (fixed_cap) {
        // every tile owns BIN_FIXED_CAP slots: ids to registers, sorted there, records written behind the tile's first slot
        int cnt = 0;
        if (b < n_bins) {
            cnt = bin_count[b];
            const int n = min(cnt, BIN_FIXED_CAP);
            int ids[BIN_FIXED_CAP];
#pragma unroll
            for (int i = 0; i < BIN_FIXED_CAP; ++i) ids[i] = i < n ? bin_list[(long long)b * BIN_FIXED_CAP + i] : 0x7fffffff;
#pragma unroll
            for (int i = 1; i < BIN_FIXED_CAP; ++i) {
#pragma unroll
                for (int j = i; j > 0; --j) {
                    const int lo = min(ids[j - 1], ids[j]), hi = max(ids[j - 1], ids[j]);
                    ids[j - 1] = lo; ids[j] = hi;
                }
            }
            int proven = 1;
#pragma unroll
            for (int i = 0; i < BIN_FIXED_CAP; ++i) {
                if (i < n) {
                    const DevSolid &S = solids[ids[i]];
                    proven &= S.conn_proven;
                    out[(long long)b * BIN_FIXED_CAP + i] = make_bin_entry(S, ids[i], ox, oy, oz, half_ext, rad_max);
                }
            }
            if (!proven) status->need_cert = 1;
            tile_proven[b] = (unsigned char)proven;
        } else if (b == n_bins) bin_insertion_sort(global_list, 0, status->n_global);
        // the true total (also of the entries that found no slot): sizes the scan + fill path if it has to take over
        unsigned tot = (unsigned)cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&status->bin_total, (int)tot);
        return;
    }
    if (b == n_bins) { bin_insertion_sort(global_list, 0, status->n_global); status->bin_total = bin_off[n_bins]; return; }
    if (b > n_bins) return;
    const int beg = bin_off[b], end = bin_off[b + 1];
    if (end > bin_cap) return;
    bin_insertion_sort(bin_list, beg, end);
    int proven = 1;   // every candidate of this tile has a provably connected cell set: its cells need no certificate pass
    for (int pos = beg; pos < end; ++pos) {
        const int s = bin_list[pos];
        const DevSolid &S = solids[s];
        proven &= S.conn_proven;
        if (!S.conn_proven) status->need_cert = 1;
        out[pos] = make_bin_entry(S, s, ox, oy, oz, half_ext, rad_max);
    }
    tile_proven[b] = (unsigned char)proven;
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
    int K = 3;
    // mesh
    bool has_mesh = false;
    DevMesh dm{};
    DevBuf<double> points, cc, V, Cf, Sf;
    DevBuf<int> cp_off, cp, cf_off, cf, fp_off, fp, nb_off, nb;
    DevBuf<float2> cell_rad;
    DevBuf<float4> cell_box;
    DevBuf<unsigned char> tile_proven;   // per tile and step: all its candidate solids are provably connected
    DevBuf<double> magSf;
    DevBuf<double2> face_rec;
    DevBuf<unsigned> hex_topo, tile_key, btopo;
    DevBuf<double> box6, cfa6;
    bool allow_box = true;      // SDFIBM_BOX=0: keep exact-box meshes on the general hexahedron kernel
    bool allow_order_free = false;   // meshes whose cells differ in vertex count: accept the order-free ALL_INSIDE rule (sdfibm_allow_order_free)
    DevBuf<int> nb6;
    DevBuf<int> orig, inv;      // tile-order renumbering: position -> caller's label and back
    DevBuf<double> cc_orig;     // cell centres in the caller's order (fixInternal)
    DevBuf<float4> cc32;
    static const int MAX_CHUNK = 64;
    int n_chunk = 8;
    int n_slab = 4;             // ... and the kernels run slab by slab (n_chunk / n_slab chunks each)           // host-buffer pipeline: U arrives / the fields leave in this many cell chunks
    int chunk_cmin[MAX_CHUNK] = {0}, chunk_cmax[MAX_CHUNK] = {0};   // caller-label range of every position chunk
    double half_ext = 0.0;
    double bmin[3], bmax[3];
    float rad3_max = 0.f, radxy_max = 0.f;
    // shapes
    std::vector<DevShape> h_shapes;
    DevBuf<DevShape> shapes;
    std::vector<sdfibm_sdf_op_t> h_ops;   // op table of the composed shapes (SDFIBM_SHAPE_PROGRAM)
    DevBuf<sdfibm_sdf_op_t> sdf_ops;
    // per step
    DevBuf<sdfibm_solid_t> solids_in;
    DevBuf<DevSolid> solids;
    DevBuf<int> bin_off, bin_list, global_list, slots;
    DevBuf<unsigned char> zero_block;   // [StepStatus | root_count | pair_counts | bin_count | bin_cursor]: zeroed by one memset per step
    StepStatus *status = nullptr;       // the pointers below live in zero_block
    int *root_count = nullptr, *bin_count = nullptr, *bin_cursor = nullptr;
    unsigned *pair_counts = nullptr;
    DevBuf<double> scal;                // {1/dt, rhof} of the step
    DevBuf<unsigned long long> slab_start;   // queue length before the current slab (host-buffer path)
    double *h_scal = nullptr;           // pinned
    DevBuf<BinEntry> bin_entries;
    DevBuf<double2> heavy_res;
    DevBuf<unsigned char> n_item;
    DevBuf<int2> heavy;
    DevBuf<double> ft_internal;
    DevBuf<unsigned char> scan_tmp;
    StepStatus *h_status = nullptr; // pinned
    sdfibm_solid_t *h_solids = nullptr; // pinned staging
    size_t h_solids_cap = 0;
    BinGrid grid{};
    int n_solids_last = 0;
    // fields kept on device for the host-buffer API and fixInternal
    DevBuf<double> dU, dAs, dFs, dTs, dCt, dFT;
    DevBuf<double> sU, sOut, sFT; // mean-field sampler scratch
    const double *last_Ct = nullptr;
    const double *last_As = nullptr, *last_Fs = nullptr, *last_Ts = nullptr;   // device outputs of the last interact (apply_forcing / touched download)
    DevBuf<int> t_flag, t_off, t_cells;      // touched-cell compaction
    DevBuf<double> t_vals;
    // replay
    DevBuf<int> labels, seed_cell, min_label, chosen, changed, flagged_list;
    DevBuf<unsigned long long> seed_key;
    DevBuf<unsigned char> excluded;
    // collision step: buffers kept between calls (evolve runs it 20 times per step; cudaMalloc / cudaFree per call cost more than the kernels)
    DevBuf<int> col_ids, col_sids, col_pcnt, col_poff, col_pairs;
    DevBuf<unsigned> col_keys, col_skeys;
    DevBuf<unsigned char> col_tmp;
    DevBuf<double> col_ft;
    cudaEvent_t ev_aux[2] = {nullptr, nullptr};   // device time of the last fixInternal / collision kernels
    double t_fix_ms = 0, t_col_ms = 0;
    bool last_used_replay = false;
    // stats
    StepStatus last{};
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // host-buffer entry: U arrives and the fields leave in cell chunks on two copy streams, overlapped with the kernels
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[MAX_CHUNK] = {}, ev_fin[MAX_CHUNK] = {};
    struct { bool active = false, stale = false; double *As = nullptr, *Fs = nullptr, *Ts = nullptr, *Ct = nullptr; const double *U = nullptr; } pipe;
    double t_host_us[4] = {0, 0, 0, 0}; // host wall time of the last interact: solid staging, enqueue / graph launch, wait for the GPU, whole call
    bool t_pending = false;              // ev[] hold a pass whose times have not been folded into t_ms yet
    double t_add = 0.0;
    double t_ms[6] = {0, 0, 0, 0, 0, 0}; // binning, k_classify, k_heavy, k_accumulate, connectivity+finalise, whole pipeline
    int n_sm = 148;
    int64_t launches = 0, graph_launches = 0;
    // the device-resident entry replays one captured graph per step while its arguments stay the same
    typedef uint64_t GraphKey[20];
    GraphKey graph_key = {0};
    cudaGraphExec_t graph_exec = nullptr;
    bool use_graph = true;
    bool bin_fixed = false;              // tile bins with BIN_FIXED_CAP slots each (set per mesh; a tile that overflows switches the context to scan + fill)
    bool bin_fixed_allowed = true;       // SDFIBM_BIN_FIXED=0
    bool part_shapes = true;             // SDFIBM_PART_SHAPES=0: one queue order for all shapes
    bool shapes_sphere_and_other = false; // the shape table holds spheres AND other shapes: only then does the partition pay
    bool classify4 = true;               // k_classify4 (four positions per thread); SDFIBM_CLASSIFY4=0: the one-position kernel
    bool shapes_may_be_global = false;   // the shape table holds a plane or a 2-D shape
    bool shapes_refinable = false;       // ... or a convex analytic shape the fp32 corner refinement of k_classify handles
    const sdfibm_solid_t *ext_solids = nullptr;   // device-resident solid records supplied by the caller for the current call
    int n_global_hint = 0;   // host-side: some solid may be on the global list (selects the k_classify variant)
    // cross-rank exchange (sdfibm_comm_*): one NCCL communicator over the ranks that share the replicated solid cloud
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_n = 1;
    bool comm_auto_reduce = false;       // every interact all-reduces the per-solid (F, T) sums before it returns
    bool comm_gather_solids = false;     // the replicated solid states arrive as one 1/N PCIe upload per rank + an all-gather
    DevBuf<double> ft_partial;           // this rank's sums (the all-reduce is out of place: a local retry keeps them)
    DevBuf<double> retry_flag;           // [0] this rank must re-run the step (overflow / replay), [1] sum over ranks
    double *h_retry_sum = nullptr;       // pinned
    DevBuf<sdfibm_solid_t> solids_slice, solids_gathered;
    double t_comm_ms = 0;                // device time of the last step's collectives
    bool gathered_now = false;           // the current call's solid records are in solids_gathered
    double *reduce_out = nullptr;        // run_pipeline: all-reduce the sums into this buffer right behind the first pass
    cudaEvent_t ev_comm[2] = {nullptr, nullptr};
    cudaStream_t s_comm = nullptr;       // high-priority stream of the split step's all-reduce
    cudaEvent_t ev_ft = nullptr;
    bool comm_split = true;              // SDFIBM_COMM_SPLIT=0: the all-reduce behind the whole step, on the context stream
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
    case SDFIBM_SHAPE_PROGRAM:   // the record carries its certified radii (include/sdfibm_b200.h)
        kind = p[4] != 0.0 ? KIND_2D : KIND_3D;
        ro = p[2];
        ri = p[3];
        break;
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


// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time: the soname every NCCL 2.x ships (a process that already holds one — torch's bundled copy —
// gets that one back, so there are never two NCCL instances in one process)
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};
static NcclApi *nccl_api() {
    static NcclApi api;
    if (api.handle || !api.error.empty()) return &api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.error = std::string("NCCL is not loadable: ") + dlerror(); return &api; }
    auto sym = [&](const char *n) { void *f = dlsym(h, n); if (!f && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n; return f; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (api.error.empty()) api.handle = h;
    return &api;
}
#define NCCL_TRY(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r__ = (expr);                                                                       \
        if (r__ != ncclSuccess)                                                                          \
            return fail(SDFIBM_ERR_CUDA, std::string(#expr) + ": " + nccl_api()->GetErrorString(r__));   \
    } while (0)

// this rank has to run the step again (a capacity grew, or a solid needs the flood-fill replay): the other ranks must learn it,
// because the all-reduce that follows is collective
__global__ void k_retry_flag(const StepStatus *st, long long heavy_cap, int global_hint, double *flag) {
    const bool again = (st->n_global > 0 && !global_hint) || st->bin_overflow || st->slot_overflow || st->heavy_total + st->heavy_gen > (unsigned long long)heavy_cap || st->n_flagged > 0;
    flag[0] = again ? 1.0 : 0.0;
}

// per-solid (F, T): sum over ranks of `partial` into `out` (one ncclAllReduce of 6N fp64 instead of the reference's 2N
// Foam::reduce calls, src/solidcloud.cpp:427-431) and, grouped into the same launch, the sum of the ranks' retry flags
static int enqueue_comm_reduce(sdfibm_context *ctx, const double *partial, double *out, int n_solids, bool with_flag) {
    NcclApi *a = nccl_api();
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaEventRecord(ctx->ev_comm[0], st));
    if (with_flag) {
        k_retry_flag<<<1, 1, 0, st>>>(ctx->status, (long long)ctx->heavy.n, ctx->n_global_hint, ctx->retry_flag.p);
        NCCL_TRY(a->GroupStart());
    }
    NCCL_TRY(a->AllReduce(partial, out, 6 * (size_t)n_solids, ncclDouble, ncclSum, ctx->comm, st));
    if (with_flag) {
        NCCL_TRY(a->AllReduce(ctx->retry_flag.p, ctx->retry_flag.p + 1, 1, ncclDouble, ncclSum, ctx->comm, st));
        NCCL_TRY(a->GroupEnd());
        CUDA_TRY(cudaMemcpyAsync(ctx->h_retry_sum, ctx->retry_flag.p + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_comm[1], st));
    return SDFIBM_OK;
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
    CUDA_TRY(cudaMallocHost(&ctx->h_scal, 4 * sizeof(double)));   // [0..1] per-step scalars, [2..3] small read-back scratch
    CUDA_TRY(cudaMallocHost(&ctx->h_retry_sum, sizeof(double)));
    *ctx->h_retry_sum = 0.0;
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreate(&ctx->ev_comm[i]));
    if (!ctx->s_comm) {
        int lo_p = 0, hi_p = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
        CUDA_TRY(cudaStreamCreateWithPriority(&ctx->s_comm, cudaStreamNonBlocking, hi_p));   // its kernels are dispatched ahead of the certificate pass
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_ft, cudaEventDisableTiming));
    }
    if (const char *e = getenv("SDFIBM_COMM_SPLIT")) ctx->comm_split = atoi(e) != 0;
    CUDA_TRY(ctx->scal.ensure(2));
    if (const char *e = getenv("SDFIBM_GRAPH")) ctx->use_graph = atoi(e) != 0;
    if (const char *e = getenv("SDFIBM_CLASSIFY4")) ctx->classify4 = atoi(e) != 0;
    if (const char *e = getenv("SDFIBM_PART_SHAPES")) ctx->part_shapes = atoi(e) != 0;
    if (const char *e = getenv("SDFIBM_BIN_FIXED")) ctx->bin_fixed_allowed = atoi(e) != 0;
    if (const char *e = getenv("SDFIBM_BOX")) ctx->allow_box = atoi(e) != 0;
    if (const char *e = getenv("SDFIBM_ALLOW_ORDER_FREE")) ctx->allow_order_free = atoi(e) != 0;
    for (int i = 0; i < 6; ++i) CUDA_TRY(cudaEventCreate(&ctx->ev[i]));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    if (const char *e = getenv("SDFIBM_SLABS")) ctx->n_slab = std::max(atoi(e), 1);
    if (const char *e = getenv("SDFIBM_CHUNKS")) ctx->n_chunk = std::min(std::max(atoi(e), 1), (int)sdfibm_context::MAX_CHUNK);
    for (int i = 0; i < sdfibm_context::MAX_CHUNK; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fin[i], cudaEventDisableTiming));
    }
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
    ctx->tile_key.release(); ctx->orig.release(); ctx->inv.release(); ctx->cc_orig.release(); ctx->cc32.release();
    ctx->tile_proven.release(); ctx->cell_box.release(); ctx->cell_rad.release(); ctx->magSf.release(); ctx->face_rec.release(); ctx->hex_topo.release(); ctx->btopo.release(); ctx->box6.release(); ctx->cfa6.release(); ctx->nb6.release(); ctx->shapes.release(); ctx->sdf_ops.release(); ctx->solids_in.release(); ctx->solids.release();
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->comm && nccl_api()->handle) nccl_api()->CommDestroy(ctx->comm);
    ctx->ft_partial.release(); ctx->retry_flag.release(); ctx->solids_slice.release(); ctx->solids_gathered.release();
    if (ctx->h_retry_sum) cudaFreeHost(ctx->h_retry_sum);
    for (int i = 0; i < 2; ++i) if (ctx->ev_comm[i]) cudaEventDestroy(ctx->ev_comm[i]);
    if (ctx->ev_ft) cudaEventDestroy(ctx->ev_ft);
    if (ctx->s_comm) cudaStreamDestroy(ctx->s_comm);
    ctx->bin_off.release(); ctx->bin_list.release(); ctx->zero_block.release(); ctx->scal.release(); ctx->slab_start.release();
    ctx->global_list.release(); ctx->slots.release();
    ctx->bin_entries.release(); ctx->heavy_res.release(); ctx->n_item.release(); ctx->heavy.release();
    ctx->ft_internal.release(); ctx->scan_tmp.release(); ctx->t_flag.release(); ctx->t_off.release(); ctx->t_cells.release(); ctx->t_vals.release();
    ctx->dU.release(); ctx->dAs.release(); ctx->dFs.release(); ctx->dTs.release(); ctx->dCt.release(); ctx->dFT.release(); ctx->sU.release(); ctx->sOut.release(); ctx->sFT.release();
    ctx->labels.release(); ctx->seed_cell.release(); ctx->min_label.release(); ctx->chosen.release();
    ctx->changed.release(); ctx->seed_key.release(); ctx->excluded.release(); ctx->flagged_list.release();
    ctx->col_ids.release(); ctx->col_sids.release(); ctx->col_pcnt.release(); ctx->col_poff.release(); ctx->col_pairs.release();
    ctx->col_keys.release(); ctx->col_skeys.release(); ctx->col_tmp.release(); ctx->col_ft.release();
    for (int i = 0; i < 2; ++i) if (ctx->ev_aux[i]) cudaEventDestroy(ctx->ev_aux[i]);
    for (int i = 0; i < 6; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < sdfibm_context::MAX_CHUNK; ++i) { if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]); if (ctx->ev_fin[i]) cudaEventDestroy(ctx->ev_fin[i]); }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->h_solids) cudaFreeHost(ctx->h_solids);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SDFIBM_OK;
}

int sdfibm_alloc_pinned(size_t bytes, void **out) {
    if (!out) return fail(SDFIBM_ERR_ARG, "sdfibm_alloc_pinned: null out");
    CUDA_TRY(cudaMallocHost(out, std::max<size_t>(bytes, 1)));
    return SDFIBM_OK;
}
int sdfibm_free_pinned(void *p) {
    if (p) CUDA_TRY(cudaFreeHost(p));
    return SDFIBM_OK;
}

static void drop_graph(sdfibm_context *ctx);

int sdfibm_set_cell_slots(sdfibm_context *ctx, int slots) {
    if (!ctx || slots < 1 || slots > 64) return fail(SDFIBM_ERR_ARG, "sdfibm_set_cell_slots: slots must be in 1..64");
    if (ctx->has_mesh) return fail(SDFIBM_ERR_STATE, "sdfibm_set_cell_slots: call before sdfibm_set_mesh");
    ctx->K = slots;
    return SDFIBM_OK;
}

int sdfibm_allow_order_free(sdfibm_context *ctx, int on) {
    if (!ctx) return fail(SDFIBM_ERR_ARG, "sdfibm_allow_order_free: null context");
    ctx->allow_order_free = on != 0;
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
    {
        // the reference calls a cell ALL_INSIDE when its inside-vertex count equals the vertex count of the cell that
        // DISCOVERED it in the flood fill (src/cellenumerator.cpp:25 compares with m_c2p[icur], the current cell) — on a mesh whose
        // cells differ in vertex count (hanging-node refinement, hex / prism / polyhedron mixes) the outcome depends on the
        // fill's visiting order.  This library uses the cell's own vertex count: the same thing when all cells have equally
        // many vertices, an order-free rule otherwise — which the caller has to accept explicitly.
        const int nv0 = m->cell_points_off[1] - m->cell_points_off[0];
        bool same = true;
        for (int c = 1; c < m->n_cells && same; ++c) same = (m->cell_points_off[c + 1] - m->cell_points_off[c]) == nv0;
        if (!same && !ctx->allow_order_free)
            return fail(SDFIBM_ERR_UNSUPPORTED,
                        "sdfibm_set_mesh: the cells differ in vertex count: the reference's ALL_INSIDE test then depends on its flood fill's "
                        "visiting order (cellenumerator.cpp:25), which is not reproduced; call sdfibm_allow_order_free(ctx, 1) "
                        "(or set SDFIBM_ALLOW_ORDER_FREE=1) to accept the order-free rule (a cell is ALL_INSIDE when all of ITS vertices are inside)");
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    drop_graph(ctx);   // a captured step holds the old mesh's pointers and constants by value
    ctx->has_mesh = false;
    const size_t nC = m->n_cells, nP = m->n_points, nF = m->n_faces;
    int rc;
    // ---- the tile grid: ~256 cells per tile (8x8x4 mean cell sizes; 16x16 columns for one-cell-thick 2-D meshes) ----
    {
        BinGrid &g = ctx->grid;
        double ext[3], T[3];
        for (int k = 0; k < 3; ++k) ext[k] = std::max(m->bounds_max[k] - m->bounds_min[k], 1e-300);
        if (two_d) {
            const double h = std::sqrt(ext[0] * ext[1] / (double)nC);
            T[0] = T[1] = 16.0 * h; T[2] = 2.0 * ext[2];
        } else {
            const double h = std::cbrt(ext[0] * ext[1] * ext[2] / (double)nC);
            T[0] = T[1] = 8.0 * h; T[2] = 4.0 * h;
        }
        if (const char *e = getenv("SDFIBM_TILE")) {   // experiments: tile extents in mean cell sizes, "tx,ty,tz"
            double t3[3];
            if (sscanf(e, "%lf,%lf,%lf", &t3[0], &t3[1], &t3[2]) == 3 && t3[0] > 0 && t3[1] > 0 && t3[2] > 0) {
                const double h = two_d ? std::sqrt(ext[0] * ext[1] / (double)nC) : std::cbrt(ext[0] * ext[1] * ext[2] / (double)nC);
                T[0] = t3[0] * h; T[1] = t3[1] * h; if (!two_d) T[2] = t3[2] * h;
            }
        }
        for (;;) {
            double nt = 1;
            for (int k = 0; k < 3; ++k) nt *= std::max(1.0, std::ceil(ext[k] / T[k]));
            if (nt <= 6.4e7) break;
            for (int k = 0; k < 3; ++k) T[k] *= 1.26;
        }
        g.n_bins = 1;
        ctx->half_ext = 0.0;
        for (int k = 0; k < 3; ++k) {
            g.lo[k] = m->bounds_min[k];
            g.inv[k] = 1.0 / T[k];
            g.n[k] = (int)std::max(1.0, std::ceil(ext[k] / T[k]));
            g.n_bins *= g.n[k];
            ctx->half_ext = std::max(ctx->half_ext, 0.5 * ext[k]);
        }
        // fixed-capacity tile bins (40 bytes per slot: id + record) as long as they stay below 1 GiB
        ctx->bin_fixed = ctx->bin_fixed_allowed && (double)g.n_bins * BIN_FIXED_CAP * 40.0 <= 1073741824.0;
    }
    DevMesh &d = ctx->dm;
    for (int k = 0; k < 3; ++k) d.origin[k] = 0.5 * (m->bounds_min[k] + m->bounds_max[k]);
    if ((rc = upload(ctx->points, m->points, 3 * nP, st))) return rc;
    if ((rc = upload(ctx->Cf, m->face_centres, 3 * nF, st))) return rc;
    if ((rc = upload(ctx->Sf, m->face_areas, 3 * nF, st))) return rc;
    if ((rc = upload(ctx->fp_off, m->face_points_off, nF + 1, st))) return rc;
    if ((rc = upload(ctx->fp, m->face_points, (size_t)m->face_points_off[nF], st))) return rc;
    if ((rc = upload(ctx->cc_orig, m->cell_centres, 3 * nC, st))) return rc;
    // ---- renumber the cells into tile order: stable sort of the cell labels by tile key ----
    CUDA_TRY(ctx->orig.ensure(nC));
    CUDA_TRY(ctx->inv.ensure(nC));
    CUDA_TRY(ctx->tile_key.ensure(nC));
    {
        DevBuf<unsigned> keys;
        DevBuf<int> ids;
        DevBuf<unsigned char> tmp;
        CUDA_TRY(keys.ensure(nC));
        CUDA_TRY(ids.ensure(nC));
        k_tile_keys<<<grid_for(nC, 256), 256, 0, st>>>(ctx->cc_orig.p, (int)nC, ctx->grid, keys.p, ids.p);
        size_t sb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sb, keys.p, ctx->tile_key.p, ids.p, ctx->orig.p, (int)nC, 0, 32, st);
        CUDA_TRY(tmp.ensure(sb));
        cub::DeviceRadixSort::SortPairs(tmp.p, sb, keys.p, ctx->tile_key.p, ids.p, ctx->orig.p, (int)nC, 0, 32, st);
        k_perm_inverse<<<grid_for(nC, 256), 256, 0, st>>>(ctx->orig.p, (int)nC, ctx->inv.p);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        keys.release(); ids.release(); tmp.release();
    }
    {
        // per-cell arrays and CSR connectivity, permuted; cellCells values are mapped to positions
        DevBuf<double> o_V;
        DevBuf<int> o_off, o_val, cnt;
        DevBuf<unsigned char> tmp;
        if ((rc = upload(o_V, m->cell_volumes, nC, st))) return rc;
        CUDA_TRY(ctx->cc.ensure(3 * nC));
        CUDA_TRY(ctx->V.ensure(nC));
        k_perm_cells<<<grid_for(nC, 256), 256, 0, st>>>(ctx->cc_orig.p, o_V.p, ctx->orig.p, (int)nC, ctx->cc.p, ctx->V.p);
        CUDA_TRY(cnt.ensure(nC + 1));
        struct Csr { const int32_t *off, *val; DevBuf<int> *d_off, *d_val; const int *map; };
        Csr csr[3] = {{m->cell_points_off, m->cell_points, &ctx->cp_off, &ctx->cp, nullptr},
                      {m->cell_faces_off, m->cell_faces, &ctx->cf_off, &ctx->cf, nullptr},
                      {m->cell_cells_off, m->cell_cells, &ctx->nb_off, &ctx->nb, ctx->inv.p}};
        for (auto &a : csr) {
            const size_t nv = (size_t)a.off[nC];
            if ((rc = upload(o_off, a.off, nC + 1, st))) return rc;
            if ((rc = upload(o_val, a.val, nv, st))) return rc;
            CUDA_TRY(a.d_off->ensure(nC + 1));
            CUDA_TRY(a.d_val->ensure(nv));
            k_perm_counts<<<grid_for(nC + 1, 256), 256, 0, st>>>(o_off.p, ctx->orig.p, (int)nC, cnt.p);
            size_t tb = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, a.d_off->p, (int)nC + 1, st);
            CUDA_TRY(tmp.ensure(tb));
            cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, a.d_off->p, (int)nC + 1, st);
            k_perm_csr<<<grid_for(nC, 256), 256, 0, st>>>(o_off.p, o_val.p, ctx->orig.p, a.d_off->p, a.map, (int)nC, a.d_val->p);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(st));   // the host arrays of this pass are consumed; o_off / o_val are reused
        }
        o_V.release(); o_off.release(); o_val.release(); cnt.release(); tmp.release();
    }
    {
        DevBuf<int> rng;
        CUDA_TRY(rng.ensure(2 * sdfibm_context::MAX_CHUNK));
        const int MC = sdfibm_context::MAX_CHUNK;
        int init[2 * sdfibm_context::MAX_CHUNK];
        for (int k = 0; k < MC; ++k) { init[k] = 0x7fffffff; init[MC + k] = -1; }
        CUDA_TRY(cudaMemcpyAsync(rng.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        k_chunk_ranges<<<grid_for(nC, 256), 256, 0, st>>>(ctx->orig.p, (int)nC, ctx->n_chunk, rng.p, rng.p + MC);
        CUDA_TRY(cudaMemcpyAsync(init, rng.p, sizeof(init), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int k = 0; k < MC; ++k) { ctx->chunk_cmin[k] = init[k]; ctx->chunk_cmax[k] = init[MC + k]; }
        rng.release();
    }
    CUDA_TRY(ctx->cell_rad.ensure(nC));
    CUDA_TRY(ctx->cc32.ensure(nC));
    CUDA_TRY(ctx->slots.ensure(nC * ctx->K));
    CUDA_TRY(ctx->n_item.ensure(nC));
    {
        size_t cap = std::max<size_t>(1 << 20, nC / 2);
        cap = std::min<size_t>(cap, (1u << 28) - 1);   // a slot record holds the queue index in 28 bits
        if (const char *e = getenv("SDFIBM_HEAVY_CAP0")) cap = std::max<size_t>(16, (size_t)atoll(e));   // tests: force the growth path
        CUDA_TRY(ctx->heavy.ensure(cap));
        CUDA_TRY(ctx->heavy_res.ensure(cap));
    }
    d.n_cells = m->n_cells; d.n_points = m->n_points; d.n_faces = m->n_faces;
    d.points = ctx->points.p; d.cc = ctx->cc.p; d.V = ctx->V.p; d.Cf = ctx->Cf.p; d.Sf = ctx->Sf.p;
    d.cp_off = ctx->cp_off.p; d.cp = ctx->cp.p; d.cf_off = ctx->cf_off.p; d.cf = ctx->cf.p;
    d.fp_off = ctx->fp_off.p; d.fp = ctx->fp.p; d.nb_off = ctx->nb_off.p; d.nb = ctx->nb.p;
    d.cell_rad = ctx->cell_rad.p;
    d.orig = ctx->orig.p; d.cc32 = ctx->cc32.p; d.tile_key = ctx->tile_key.p;
    d.two_d = two_d ? 1 : 0;
    // hexahedral fast path?
    bool is_hex = true;
    for (size_t c = 0; c < nC && is_hex; ++c)
        is_hex = (m->cell_points_off[c + 1] - m->cell_points_off[c] == 8) && (m->cell_faces_off[c + 1] - m->cell_faces_off[c] == 6);
    for (size_t f = 0; f < nF && is_hex; ++f) is_hex = (m->face_points_off[f + 1] - m->face_points_off[f] == 4);
    is_hex = is_hex && m->cell_points_off[0] == 0 && m->cell_faces_off[0] == 0 && m->face_points_off[0] == 0;
    d.is_hex = is_hex ? 1 : 0;
    size_t n_hex_cells = 0;
    if (!is_hex && m->cell_points_off[0] == 0 && m->cell_faces_off[0] == 0 && m->face_points_off[0] == 0) {
        for (size_t c = 0; c < nC; ++c) {
            bool h = (m->cell_points_off[c + 1] - m->cell_points_off[c] == 8) && (m->cell_faces_off[c + 1] - m->cell_faces_off[c] == 6);
            for (int k = m->cell_faces_off[c]; k < m->cell_faces_off[c + 1] && h; ++k) {
                const int f = m->cell_faces[k];
                h = m->face_points_off[f + 1] - m->face_points_off[f] == 4;
            }
            n_hex_cells += h;
        }
    }
    d.mixed = (!is_hex && n_hex_cells > 0) ? 1 : 0;
    const bool hex_path = is_hex || d.mixed;
    CUDA_TRY(ctx->magSf.ensure(nF));
    d.magSf = ctx->magSf.p;
    d.hex_topo = nullptr;
    d.face_rec = nullptr;
    if (hex_path) {
        CUDA_TRY(ctx->face_rec.ensure(4 * nF));
        d.face_rec = ctx->face_rec.p;
    }
    k_face_mag<<<grid_for(nF, 256), 256, 0, st>>>(d.Cf, d.Sf, (int)nF, ctx->magSf.p, ctx->face_rec.p);
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
    if (hex_path) {
        CUDA_TRY(ctx->hex_topo.ensure(3 * nC));
        d.hex_topo = ctx->hex_topo.p;
        k_hex_topo<<<grid_for(nC, 256), 256, 0, st>>>(d, ctx->hex_topo.p, bad.p);
    }
    // exact-box fast path of the exact evaluation (k_heavy_box)?
    d.box_exact = 0; d.box6 = nullptr; d.btopo = nullptr; d.cfa6 = nullptr;
    if (is_hex && ctx->allow_box) {
        DevBuf<int> okf;
        CUDA_TRY(okf.ensure(2));
        const int one[2] = {1, 1};
        CUDA_TRY(cudaMemcpyAsync(okf.p, one, sizeof(one), cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx->box6.ensure(6 * nC));
        CUDA_TRY(ctx->btopo.ensure(2 * nC));
        CUDA_TRY(ctx->cfa6.ensure(6 * nC));
        k_box_topo<<<grid_for(nC, 128), 128, 0, st>>>(d, ctx->box6.p, ctx->btopo.p, ctx->cfa6.p, okf.p);
        int h_ok[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(h_ok, okf.p, sizeof(h_ok), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        okf.release();
        if (h_ok[0]) {
            d.box_exact = h_ok[1] ? 2 : 1;
            d.box6 = ctx->box6.p; d.btopo = ctx->btopo.p;
            if (h_ok[1]) ctx->cfa6.release();
            d.cfa6 = ctx->cfa6.p;
        } else { ctx->box6.release(); ctx->btopo.release(); ctx->cfa6.release(); }
    }
    CUDA_TRY(ctx->nb6.ensure(6 * nC));
    d.nb6 = ctx->nb6.p;
    k_nb6<<<grid_for(nC, 256), 256, 0, st>>>(d, ctx->nb6.p);
    {
        // identical cell volumes (uniform lattices): one constant
        DevBuf<int> differs;
        CUDA_TRY(differs.ensure(1));
        CUDA_TRY(cudaMemsetAsync(differs.p, 0, sizeof(int), st));
        k_v_uniform<<<grid_for(nC, 256), 256, 0, st>>>(d.V, (int)nC, differs.p);
        int h_differs = 1;
        double h_v0 = 0.0;
        CUDA_TRY(cudaMemcpyAsync(&h_differs, differs.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&h_v0, d.V, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        differs.release();
        d.V_uniform = h_differs ? 0 : 1;
        d.V_const = h_v0;
    }
    {
        // cell boxes; a mesh of identical boxes keeps one constant instead of 16 bytes per cell
        DevBuf<int> ext;
        CUDA_TRY(ctx->cell_box.ensure(nC));
        CUDA_TRY(ext.ensure(7));
        const int init[7] = {0, 0, 0, 0x7f7fffff, 0x7f7fffff, 0x7f7fffff, 0};
        CUDA_TRY(cudaMemcpyAsync(ext.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        k_cell_box<<<grid_for(nC, 256), 256, 0, st>>>(d, ctx->cell_box.p, ext.p);
        int h_ext[7];
        CUDA_TRY(cudaMemcpyAsync(h_ext, ext.p, sizeof(h_ext), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        ext.release();
        float mx[3], mn[3];
        memcpy(mx, h_ext, sizeof(mx));
        memcpy(mn, h_ext + 3, sizeof(mn));
        bool uniform = h_ext[6] == 0;
        for (int k = 0; k < 3; ++k) uniform = uniform && (mx[k] - mn[k] <= 1e-6f * mx[k]);
        d.box_uniform = uniform ? 1 : 0;
        d.box_const = make_float4(mx[0], mx[1], mx[2], 1.f);
        d.lattice_full = 0;
        if (uniform) {
            long long n[3];
            bool ok = true;
            for (int k = 0; k < 3; ++k) {
                const double sz = 2.0 * (double)mx[k] / (1.0 + REL_MARGIN), ext = m->bounds_max[k] - m->bounds_min[k];
                n[k] = std::llround(ext / sz);
                ok = ok && n[k] >= 1 && std::fabs((double)n[k] * sz - ext) <= 1e-4 * sz;
                d.lattice_lo[k] = m->bounds_min[k]; d.lattice_hi[k] = m->bounds_max[k];
            }
            ok = ok && n[0] * n[1] * n[2] == (long long)nC &&
                 (long long)m->n_internal_faces == 3 * n[0] * n[1] * n[2] - (n[0] * n[1] + n[1] * n[2] + n[0] * n[2]);
            d.lattice_full = ok ? 1 : 0;
        }
        d.cell_box = ctx->cell_box.p;
        if (uniform) { ctx->cell_box.release(); d.cell_box = nullptr; }
    }
    k_cc32<<<grid_for(nC, 256), 256, 0, st>>>(ctx->cc.p, ctx->cell_rad.p, (int)nC, d.origin[0], d.origin[1], d.origin[2], ctx->cc32.p);
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

// stack discipline of a composed shape's program: every op finds its operands, depths stay within SDFIBM_SDF_STACK, exactly
// one value is left
static std::string check_sdf_program(const std::vector<sdfibm_sdf_op_t> &ops, const sdfibm_shape_t &rec) {
    const double first = rec.p[0], count = rec.p[1];
    if (!(first >= 0) || !(count >= 1) || first != std::floor(first) || count != std::floor(count) || first + count > (double)ops.size())
        return "program range outside the op table (call sdfibm_set_shape_programs first)";
    if (!(rec.p[2] > 0) || !(rec.p[3] >= 0) || rec.p[3] > rec.p[2]) return "a composed shape needs certified radii: p[2] = outer > 0, 0 <= p[3] = inner <= outer";
    int np = 0, nv = 0;
    for (int i = (int)first; i < (int)(first + count); ++i) {
        const int op = ops[i].op;
        if (op == SDFIBM_OP_POINT || op == SDFIBM_OP_POINT_2D) ++np;
        else if (op >= SDFIBM_OP_OFFSET && op <= SDFIBM_OP_FLIPY) { if (np < 1) return "a point transformation without a point"; }
        else if (op >= SDFIBM_OP_CIRCLE && op <= SDFIBM_OP_HALFSPACE) { if (np < 1) return "a primitive without a point"; --np; ++nv; }
        else if (op >= SDFIBM_OP_UNION && op <= SDFIBM_OP_DIFF) { if (nv < 2) return "a Boolean operation without two values"; --nv; }
        else return "unknown opcode " + std::to_string(op);
        if (np > SDFIBM_SDF_STACK || nv > SDFIBM_SDF_STACK) return "stack deeper than SDFIBM_SDF_STACK";
    }
    if (nv != 1 || np != 0) return "the program must consume its points and leave exactly one value";
    return "";
}

int sdfibm_set_shape_programs(sdfibm_context *ctx, const sdfibm_sdf_op_t *ops, int n_ops) {
    if (!ctx || n_ops < 0 || (n_ops > 0 && !ops)) return fail(SDFIBM_ERR_ARG, "sdfibm_set_shape_programs: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    drop_graph(ctx);
    ctx->h_ops.assign(ops, ops + n_ops);
    const int rc = upload(ctx->sdf_ops, ctx->h_ops.data(), (size_t)n_ops, ctx->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SDFIBM_OK;
}

int sdfibm_set_shapes(sdfibm_context *ctx, const sdfibm_shape_t *shapes, int n) {
    if (!ctx || !shapes || n <= 0) return fail(SDFIBM_ERR_ARG, "sdfibm_set_shapes: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    drop_graph(ctx);   // the captured step holds n_shapes and the table pointer by value
    ctx->h_shapes.resize(n);
    for (int i = 0; i < n; ++i) {
        if (shapes[i].tag < 0 || shapes[i].tag >= SDFIBM_SHAPE_NTAGS)
            return fail(SDFIBM_ERR_UNSUPPORTED, "sdfibm_set_shapes: shape type has no device tag (no CPU fallback)");
        if (shapes[i].tag == SDFIBM_SHAPE_PROGRAM) {
            const std::string err = check_sdf_program(ctx->h_ops, shapes[i]);
            if (!err.empty()) return fail(SDFIBM_ERR_ARG, "sdfibm_set_shapes: shape " + std::to_string(i) + ": " + err);
        }
        shape_bounds(shapes[i], ctx->h_shapes[i]);
    }
    ctx->shapes_may_be_global = false;
    ctx->shapes_refinable = false;
    bool any_sphere = false, any_other = false;
    for (auto &sh : ctx->h_shapes) {
        (sh.s.tag == SDFIBM_SHAPE_SPHERE ? any_sphere : any_other) = true;
        ctx->shapes_may_be_global |= (sh.kind != KIND_3D);
        ctx->shapes_refinable |= (sh.s.tag == SDFIBM_SHAPE_ELLIPSOID || sh.s.tag == SDFIBM_SHAPE_ELLIPSE || sh.s.tag == SDFIBM_SHAPE_BOX || sh.s.tag == SDFIBM_SHAPE_RECTANGLE);
    }
    ctx->shapes_sphere_and_other = any_sphere && any_other;
    if (const char *e = getenv("SDFIBM_REFINE")) ctx->shapes_refinable = ctx->shapes_refinable && atoi(e) != 0;
    int rc = upload(ctx->shapes, ctx->h_shapes.data(), (size_t)n, ctx->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SDFIBM_OK;
}

static int run_pipeline(sdfibm_context *ctx, int n_solids, const double *dU, double dt, double rhof, double *dAs,
                        double *dFs, double *dTs, double *dCt, double *dFT, bool replay);

static int stage_solids(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n, bool may_gather = false) {
    if ((size_t)n > ctx->h_solids_cap) {
        if (ctx->h_solids) cudaFreeHost(ctx->h_solids);
        ctx->h_solids = nullptr;
        ctx->h_solids_cap = 0;
        CUDA_TRY(cudaMallocHost(&ctx->h_solids, sizeof(sdfibm_solid_t) * (size_t)n));
        ctx->h_solids_cap = n;
    }
    const int ns = (int)ctx->h_shapes.size();
    // may any solid be tested by every cell (plane, or a 2-D shape whose axis is not exactly world z)?  A superset of the device's
    // decision in k_solid_prepare: it only selects the k_classify variant that merges the global list.  Shape tables without
    // planes / 2-D shapes need no per-solid pass on the host at all (the shape index is range-checked on the device).
    int hint = 0;
    if (ctx->shapes_may_be_global) {
        for (int i = 0; i < n; ++i) {
            if (solids[i].shape < 0 || solids[i].shape >= ns) return fail(SDFIBM_ERR_ARG, "solid refers to an unknown shape index");
            const int kind = ctx->h_shapes[solids[i].shape].kind;
            hint |= (kind == KIND_PLANE) || (kind == KIND_2D && (solids[i].quat[1] != 0.0 || solids[i].quat[2] != 0.0));
        }
    }
    ctx->n_global_hint = hint;
    ctx->gathered_now = false;
    if (may_gather && ctx->comm && ctx->comm_gather_solids && ctx->comm_n > 1 && n >= 8 * ctx->comm_n) {
        // the solid cloud is replicated (every rank's host holds the same array): each rank uploads one 1/N slice over PCIe and
        // the slices are all-gathered over NVLink instead of N identical full uploads
        const size_t per = ((size_t)n + ctx->comm_n - 1) / ctx->comm_n;
        CUDA_TRY(ctx->solids_slice.ensure(per));
        CUDA_TRY(ctx->solids_gathered.ensure(per * ctx->comm_n));
        const size_t b = std::min((size_t)n, per * ctx->comm_rank), e = std::min((size_t)n, b + per);
        const void *src = solids + b;
        cudaPointerAttributes attr;
        if (e > b && (cudaPointerGetAttributes(&attr, solids) != cudaSuccess || attr.type != cudaMemoryTypeHost)) {
            cudaGetLastError();
            memcpy(ctx->h_solids, solids + b, sizeof(sdfibm_solid_t) * (e - b));
            src = ctx->h_solids;
        }
        if (e > b) CUDA_TRY(cudaMemcpyAsync(ctx->solids_slice.p, src, sizeof(sdfibm_solid_t) * (e - b), cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(nccl_api()->AllGather(ctx->solids_slice.p, ctx->solids_gathered.p, per * sizeof(sdfibm_solid_t), ncclChar, ctx->comm, ctx->stream));
        ctx->gathered_now = true;
        return SDFIBM_OK;
    }
    CUDA_TRY(ctx->solids_in.ensure(n));
    // page-locked caller memory (sdfibm_alloc_pinned, cudaHostRegister, ...) is copied from directly; anything else is staged
    const void *src = solids;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, solids) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        memcpy(ctx->h_solids, solids, sizeof(sdfibm_solid_t) * (size_t)n);
        src = ctx->h_solids;
    }
    CUDA_TRY(cudaMemcpyAsync(ctx->solids_in.p, src, sizeof(sdfibm_solid_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    return SDFIBM_OK;
}

int sdfibm_interact_device(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, const double *dU, double dt,
                           double rhof, double *dAs, double *dFs, double *dTs, double *dCt, double *dFT) {
    if (!ctx || n_solids < 0 || (n_solids > 0 && ((!solids && !ctx->ext_solids) || !dFT)) || !dU || !dAs || !dFs || !dTs || !dCt)
        return fail(SDFIBM_ERR_ARG, "sdfibm_interact: null argument");
    if (!ctx->has_mesh) return fail(SDFIBM_ERR_STATE, "sdfibm_interact: set mesh and shapes first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n_solids == 0) {   // an empty cloud: the four fields are reset and nothing else happens (solidcloud.cpp:438-441)
        const size_t nC0 = ctx->dm.n_cells;
        CUDA_TRY(cudaMemsetAsync(dAs, 0, sizeof(double) * nC0, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(dFs, 0, sizeof(double) * 3 * nC0, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(dTs, 0, sizeof(double) * nC0, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(dCt, 0, sizeof(double) * nC0, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(ctx->n_item.p, 0, nC0, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->last = StepStatus{};
        ctx->last_Ct = dCt; ctx->last_As = dAs; ctx->last_Fs = dFs; ctx->last_Ts = dTs;
        ctx->n_solids_last = 0;
        return SDFIBM_OK;
    }
    if (ctx->h_shapes.empty()) return fail(SDFIBM_ERR_STATE, "sdfibm_interact: set mesh and shapes first");
    if (n_solids > (1 << 28) - 4) return fail(SDFIBM_ERR_ARG, "too many solids");
    int rc = SDFIBM_OK;
    const auto h0 = std::chrono::steady_clock::now();
    if (!ctx->pipe.active && !ctx->ext_solids) rc = stage_solids(ctx, solids, n_solids, true);   // the host-buffer entry stages them ahead of its U copies
    if (rc) return rc;
    const auto h1 = std::chrono::steady_clock::now();
    ctx->t_host_us[0] = std::chrono::duration<double, std::micro>(h1 - h0).count();
    ctx->launches = 0;
    // cross-rank sum of the per-solid (F, T) (src/solidcloud.cpp:427-431) on the context stream, right behind the kernels: the
    // step's sums go to a buffer of their own and the all-reduce writes the caller's array, so a rank that has to re-run the
    // step (queue growth, flood-fill replay) still holds its partial sums; the ranks' "again" flags ride the same NCCL launch
    const bool reduce = ctx->comm && ctx->comm_auto_reduce && ctx->comm_n > 1;
    double *const dFT_out = dFT;
    if (reduce) {
        CUDA_TRY(ctx->ft_partial.ensure(6 * (size_t)n_solids));
        CUDA_TRY(ctx->retry_flag.ensure(2));
        dFT = ctx->ft_partial.p;
        ctx->reduce_out = dFT_out;
        *ctx->h_retry_sum = 0.0;
    }
    ctx->t_comm_ms = 0;
    rc = run_pipeline(ctx, n_solids, dU, dt, rhof, dAs, dFs, dTs, dCt, dFT, false);
    ctx->reduce_out = nullptr;
    ctx->t_host_us[3] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count();
    if (rc) return rc;
    ctx->flagged_last = ctx->last.n_flagged;
    ctx->last_used_replay = false;
    cudaStream_t st = ctx->stream;
    const size_t nC = ctx->dm.n_cells;
    if (ctx->last.n_flagged > 0) {
        // exact flood-fill replay for the flagged solids
        const size_t K = ctx->K;
        CUDA_TRY(ctx->labels.ensure(nC * K));
        CUDA_TRY(ctx->excluded.ensure(nC * K));
        CUDA_TRY(ctx->seed_key.ensure(n_solids));
        CUDA_TRY(ctx->seed_cell.ensure(n_solids));
        CUDA_TRY(ctx->min_label.ensure(n_solids));
        CUDA_TRY(ctx->chosen.ensure(n_solids));
        CUDA_TRY(ctx->changed.ensure(REPLAY_BATCH));
        ReplayParams R;
        R.m = ctx->dm; R.solids = ctx->solids.p; R.n_item = ctx->n_item.p; R.slots = ctx->slots.p; R.K = ctx->K;
        R.root_count = ctx->root_count; R.labels = ctx->labels.p; R.changed = ctx->changed.p;
        R.seed_key = ctx->seed_key.p; R.seed_cell = ctx->seed_cell.p; R.min_label = ctx->min_label.p;
        R.chosen = ctx->chosen.p; R.excluded = ctx->excluded.p; R.inv = ctx->inv.p; R.grid = ctx->grid; R.bin_off = ctx->bin_off.p;
        R.bin_list = ctx->bin_list.p; R.global_list = ctx->global_list.p; R.n_global = ctx->last.n_global;
        R.n_solids = n_solids;
        const int g = grid_for(nC, 256);
        k_replay_init<<<g, 256, 0, st>>>(R);
        CUDA_TRY(cudaMemsetAsync(ctx->seed_key.p, 0xff, sizeof(unsigned long long) * n_solids, st));
        CUDA_TRY(cudaMemsetAsync(ctx->seed_cell.p, 0x7f, sizeof(int) * n_solids, st));
        CUDA_TRY(cudaMemsetAsync(ctx->min_label.p, 0x7f, sizeof(int) * n_solids, st));
        int *h_changed = reinterpret_cast<int *>(ctx->h_scal + 2);   // page-locked read-back scratch
        for (int it = 0; it < 1000000; ++it) {
            CUDA_TRY(cudaMemsetAsync(ctx->changed.p, 0, sizeof(int) * REPLAY_BATCH, st));
            for (int rep = 0; rep < REPLAY_BATCH; ++rep) k_replay_propagate<<<g, 256, 0, st>>>(R, rep);
            CUDA_TRY(cudaMemcpyAsync(h_changed, ctx->changed.p + (REPLAY_BATCH - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            ctx->launches += REPLAY_BATCH;
            if (!*h_changed) break;     // the last sweep of the batch changed nothing (or never ran: an earlier one converged)
        }
        CUDA_TRY(ctx->flagged_list.ensure((size_t)n_solids + 1));
        CUDA_TRY(cudaMemsetAsync(ctx->flagged_list.p + n_solids, 0, sizeof(int), st));
        k_replay_flagged<<<grid_for(n_solids, 256), 256, 0, st>>>(R, ctx->flagged_list.p, ctx->flagged_list.p + n_solids);
        k_replay_seed<<<g, 256, 0, st>>>(R, ctx->flagged_list.p, ctx->flagged_list.p + n_solids, 0);
        k_replay_seed<<<g, 256, 0, st>>>(R, ctx->flagged_list.p, ctx->flagged_list.p + n_solids, 1);
        k_replay_choose<<<grid_for(n_solids, 256), 256, 0, st>>>(R);
        k_replay_mark<<<g, 256, 0, st>>>(R);
        CUDA_TRY(cudaGetLastError());
        ctx->launches += 5;
        // second pass of the same kernels with the pairs outside the seed's component masked out
        rc = run_pipeline(ctx, n_solids, dU, dt, rhof, dAs, dFs, dTs, dCt, dFT, true);
        if (rc) return rc;
        ctx->last_used_replay = true;
    }
    if (reduce && *ctx->h_retry_sum > 0.0) {
        // some rank ran its step again: every rank reduces again (the flag sum is the same number everywhere)
        rc = enqueue_comm_reduce(ctx, ctx->ft_partial.p, dFT_out, n_solids, false);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    ctx->last_Ct = dCt; ctx->last_As = dAs; ctx->last_Fs = dFs; ctx->last_Ts = dTs;
    ctx->n_solids_last = n_solids;
    return SDFIBM_OK;
}

// The tail of a pass: connectivity certificate, status totals (+ the rhof scaling unless the head did it), status read-back.
static int enqueue_tail(sdfibm_context *ctx, int n_solids, double *dFT, bool replay, bool capturing, int scale) {
    cudaStream_t st = ctx->stream;
    const int nC = ctx->dm.n_cells;
    if (!replay) {
        ConnParams C;
        C.m = ctx->dm; C.solids = ctx->solids.p; C.n_item = ctx->n_item.p; C.slots = ctx->slots.p; C.K = ctx->K; C.root_count = ctx->root_count; C.tile_proven = ctx->tile_proven.p; C.status = ctx->status;
        k_connectivity<<<std::min(grid_for(nC, 256), ctx->n_sm * 8), 256, 0, st>>>(C);
        ++ctx->launches;
    }
    k_finalize<<<std::min(grid_for(n_solids, 256), 296), 256, 0, st>>>(ctx->pair_counts, ctx->root_count, n_solids, ctx->status, dFT, ctx->scal.p, scale);
    ++ctx->launches;
    CUDA_TRY(capturing ? cudaEventRecordWithFlags(ctx->ev[5], st, cudaEventRecordExternal) : cudaEventRecord(ctx->ev[5], st));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(ctx->h_status, ctx->status, sizeof(StepStatus), cudaMemcpyDeviceToHost, st));
    return SDFIBM_OK;
}

// Everything one pass of the pipeline enqueues on the context stream (directly, or once into a CUDA graph that later steps
// re-launch: ~20 dependent launches / memsets become one submission, which matters because every step starts on an idle GPU).
static int enqueue_pipeline(sdfibm_context *ctx, int n_solids, const double *dU, double *dAs, double *dFs, double *dTs, double *dCt,
                            double *dFT, bool replay, bool chunked, bool capturing, int part = 0) {
    cudaStream_t st = ctx->stream;
    const int nC = ctx->dm.n_cells;
    const BinGrid &g = ctx->grid;
    auto rec = [&](cudaEvent_t e) { return capturing ? cudaEventRecordWithFlags(e, st, cudaEventRecordExternal) : cudaEventRecord(e, st); };
    CUDA_TRY(rec(ctx->ev[0]));
    CUDA_TRY(cudaMemcpyAsync(ctx->scal.p, ctx->h_scal, 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(dFT, 0, sizeof(double) * 6 * n_solids, st));
    if (!replay) {
        // status word, root / pair counters, bin counters and cursors live in one block: one memset
        CUDA_TRY(cudaMemsetAsync(ctx->zero_block.p, 0, ctx->zero_block.n, st));
        PrepParams P;
        P.solids = ctx->ext_solids ? ctx->ext_solids : (ctx->gathered_now ? ctx->solids_gathered.p : ctx->solids_in.p); P.shapes = ctx->shapes.p; P.n_solids = n_solids; P.n_shapes = (int)ctx->h_shapes.size();
        P.out = ctx->solids.p; P.grid = g; P.rad3_max = ctx->rad3_max; P.radxy_max = ctx->radxy_max;
        for (int d = 0; d < 3; ++d) { P.mesh_lo[d] = ctx->bmin[d]; P.mesh_hi[d] = ctx->bmax[d]; P.origin[d] = ctx->dm.origin[d]; }
        P.half_ext = ctx->half_ext; P.rad_max = (double)std::max(ctx->rad3_max, ctx->radxy_max);
        P.lattice_full = ctx->dm.lattice_full;
        P.box_h[0] = ctx->dm.box_const.x; P.box_h[1] = ctx->dm.box_const.y; P.box_h[2] = ctx->dm.box_const.z;
        P.bin_count = ctx->bin_count; P.global_list = ctx->global_list.p; P.status = ctx->status;
        P.fixed_cap = ctx->bin_fixed ? BIN_FIXED_CAP : 0; P.bin_list = ctx->bin_list.p;
        k_solid_prepare<<<grid_for((long long)n_solids * BIN_SUB, 128), 128, 0, st>>>(P);
        const int bin_cap_i = (int)std::min<size_t>(ctx->bin_list.n, 0x7fffffff);
        if (!ctx->bin_fixed) {
            // scan + fill: a CSR list of exactly the size the step needs (the fallback when some tile has more candidates than slots)
            size_t tmp_bytes = ctx->scan_tmp.n;
            cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp_bytes, ctx->bin_count, ctx->bin_off.p, g.n_bins + 1, st);
            FillParams F;
            F.solids = ctx->solids.p; F.n_solids = n_solids; F.grid = g; F.rad3_max = ctx->rad3_max; F.radxy_max = ctx->radxy_max;
            for (int d = 0; d < 3; ++d) { F.mesh_lo[d] = ctx->bmin[d]; F.mesh_hi[d] = ctx->bmax[d]; }
            F.bin_off = ctx->bin_off.p; F.bin_cursor = ctx->bin_cursor; F.bin_list = ctx->bin_list.p;
            F.bin_cap = bin_cap_i; F.status = ctx->status;
            k_bin_fill<<<grid_for((long long)n_solids * BIN_SUB, 128), 128, 0, st>>>(F);
            ctx->launches += 2;
        }
        k_bin_sort_entries<<<grid_for((long long)g.n_bins + 1, 128), 128, 0, st>>>(
            ctx->bin_off.p, ctx->bin_list.p, g.n_bins, bin_cap_i, ctx->global_list.p, ctx->status, ctx->solids.p, ctx->bin_entries.p,
            ctx->dm.origin[0], ctx->dm.origin[1], ctx->dm.origin[2], ctx->half_ext, (double)std::max(ctx->rad3_max, ctx->radxy_max),
            ctx->tile_proven.p, ctx->bin_fixed ? BIN_FIXED_CAP : 0, ctx->bin_count);
        ctx->launches += 2;
    } else {
        // keep the binning of the first pass; restore the counters the status word carries
        CUDA_TRY(cudaMemsetAsync(ctx->root_count, 0, sizeof(int) * n_solids, st));
        CUDA_TRY(cudaMemsetAsync(ctx->pair_counts, 0, sizeof(unsigned) * 3 * n_solids, st));
        StepStatus keep{};
        keep.n_global = ctx->last.n_global;
        keep.bin_total = ctx->last.bin_total;
        CUDA_TRY(cudaMemcpyAsync(ctx->status, &keep, sizeof(StepStatus), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    InteractParams I;
    I.m = ctx->dm; I.solids = ctx->solids.p; I.shapes = ctx->shapes.p; I.ops = ctx->h_ops.empty() ? nullptr : ctx->sdf_ops.p; I.n_solids = n_solids; I.grid = g;
    I.bin_fixed = ctx->bin_fixed ? 1 : 0; I.bin_count = ctx->bin_count;
    // (the one-position kernel of the global-list variant queues everything in front: no partition then)
    I.part_shapes = (ctx->part_shapes && ctx->shapes_sphere_and_other && ctx->classify4 && ctx->dm.box_exact && !ctx->dm.mixed && !ctx->n_global_hint) ? 1 : 0;
    I.bin_off = ctx->bin_off.p; I.bin_list = ctx->bin_list.p; I.bin_entries = ctx->bin_entries.p; I.global_list = ctx->global_list.p; I.U = dU;
    I.scal = ctx->scal.p; I.As = dAs; I.Fs = dFs; I.Ts = dTs; I.Ct = dCt; I.force_torque = dFT;
    I.pair_counts = ctx->pair_counts; I.slots = ctx->slots.p; I.K = ctx->K;
    I.n_item = ctx->n_item.p; I.heavy = ctx->heavy.p; I.heavy_res = ctx->heavy_res.p;
    I.heavy_count = &ctx->status->heavy_total; I.heavy_gen = &ctx->status->heavy_gen; I.heavy_cap = (long long)ctx->heavy.n;
    I.excluded = replay ? ctx->excluded.p : nullptr; I.status = ctx->status;
    auto launch_classify = [&](int p0, int p1) {
        I.cls_begin = p0; I.cls_end = p1;
        const int grid = grid_for(p1 - p0, 256);
        // variants: with / without the global-list merge (planes, tilted 2-D solids) and the fp32 corner refinement (shape tables
        // holding ellipsoids, boxes, ellipses, rectangles); the plain one runs at 32 registers / full occupancy
        if (ctx->classify4 && !ctx->n_global_hint) {   // no plane / tilted 2-D solid about: four consecutive positions per thread
            const bool spec = !ctx->shapes_may_be_global && ctx->dm.box_uniform && ctx->dm.box_const.w != 0.f;   // 3-D shapes on identical box cells
            const int g4 = grid_for(p1 - p0, 4 * CLS4_NT);
            const int variant = (spec ? 4 : 0) | (ctx->shapes_refinable ? 2 : 0) | (I.part_shapes ? 1 : 0);
            switch (variant) {
                case 0: k_classify4<CLS4_NT, CLS4_MINB, false, false, false><<<g4, CLS4_NT, 0, st>>>(I); break;
                case 1: k_classify4<CLS4_NT, CLS4_MINB, false, false, true><<<g4, CLS4_NT, 0, st>>>(I); break;
                case 2: k_classify4<CLS4_NT, CLS4_MINB, false, true, false><<<g4, CLS4_NT, 0, st>>>(I); break;
                case 3: k_classify4<CLS4_NT, CLS4_MINB, false, true, true><<<g4, CLS4_NT, 0, st>>>(I); break;
                case 4: k_classify4<CLS4_NT, CLS4_MINB, true, false, false><<<g4, CLS4_NT, 0, st>>>(I); break;
                case 5: k_classify4<CLS4_NT, CLS4_MINB, true, false, true><<<g4, CLS4_NT, 0, st>>>(I); break;
                case 6: k_classify4<CLS4_NT, CLS4_MINB, true, true, false><<<g4, CLS4_NT, 0, st>>>(I); break;
                default: k_classify4<CLS4_NT, CLS4_MINB, true, true, true><<<g4, CLS4_NT, 0, st>>>(I); break;
            }
            return;
        }
        constexpr int CM = 256 / CLS_NT;   // resident CTAs scale with the CTA size: the same threads per SM
        const int gridc = grid_for(p1 - p0, CLS_NT);
        if (ctx->n_global_hint) k_classify<CLS_NT, 4 * CM, true, true><<<gridc, CLS_NT, 0, st>>>(I);
        else if (ctx->shapes_refinable) k_classify<CLS_NT, 6 * CM, false, true><<<gridc, CLS_NT, 0, st>>>(I);
        else k_classify<CLS_NT, 8 * CM, false, false><<<gridc, CLS_NT, 0, st>>>(I);
    };
    auto launch_heavy = [&]() {
        // (PROG: the shape table holds a composed shape — only then do the kernels carry the op interpreter)
        const bool prog = !ctx->h_ops.empty();
        const int gb = ctx->n_sm * BOX_CTAS_PER_SM, gh = ctx->n_sm * HEAVY_CTAS_PER_SM;
        const bool part = I.part_shapes != 0;
        if (ctx->dm.box_exact == 2) {
            if (prog) { if (part) k_heavy_box<false, true, true><<<gb, TPB, 0, st>>>(I); else k_heavy_box<false, true, false><<<gb, TPB, 0, st>>>(I); }
            else { if (part) k_heavy_box<false, false, true><<<gb, TPB, 0, st>>>(I); else k_heavy_box<false, false, false><<<gb, TPB, 0, st>>>(I); }
        } else if (ctx->dm.box_exact == 1) {
            if (prog) { if (part) k_heavy_box<true, true, true><<<gb, TPB, 0, st>>>(I); else k_heavy_box<true, true, false><<<gb, TPB, 0, st>>>(I); }
            else { if (part) k_heavy_box<true, false, true><<<gb, TPB, 0, st>>>(I); else k_heavy_box<true, false, false><<<gb, TPB, 0, st>>>(I); }
        }
        else if (ctx->dm.is_hex || ctx->dm.mixed) { if (prog) k_heavy_hex<HEAVY_CTAS_PER_SM, true><<<gh, TPB, 0, st>>>(I); else k_heavy_hex<HEAVY_CTAS_PER_SM, false><<<gh, TPB, 0, st>>>(I); }
        if (!ctx->dm.is_hex) { if (prog) k_heavy_general<true><<<gh, TPB, 0, st>>>(I); else k_heavy_general<false><<<gh, TPB, 0, st>>>(I); }
    };
    I.heavy_start = nullptr;
    CUDA_TRY(rec(ctx->ev[1]));
    if (!chunked) {
        launch_classify(0, nC);
        CUDA_TRY(rec(ctx->ev[2]));
        launch_heavy();
        CUDA_TRY(rec(ctx->ev[3]));
    }
    if (chunked) {
        CUDA_TRY(rec(ctx->ev[2]));   // (the per-kernel split is not resolved on this path: classify / heavy / final interleave)
        CUDA_TRY(rec(ctx->ev[3]));
        // Host-buffer path: the mesh is worked in N_SLAB position slabs (classify -> heavy -> final, chunk by chunk), so the first
        // output chunk leaves after a quarter of the kernel time instead of after all of it; k_final of chunk i waits for its
        // slice of U and releases its slice of the fields to the copy-out stream.  The copy engines are FIFO across streams, so
        // the U chunks are enqueued only now — after every small upload / memset the preceding kernels depend on — and still
        // start at t ~ 0 because enqueueing is asynchronous.
        for (int i = 0; i < ctx->n_chunk; ++i) {
            const size_t c0 = (size_t)nC * i / ctx->n_chunk, c1 = (size_t)nC * (i + 1) / ctx->n_chunk;
            if (c1 > c0) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(dU) + 3 * c0, ctx->pipe.U + 3 * c0, sizeof(double) * 3 * (c1 - c0), cudaMemcpyHostToDevice, ctx->s_in));
            CUDA_TRY(cudaEventRecord(ctx->ev_in[i], ctx->s_in));
        }
        // Position chunk i touches the caller's cells [chunk_cmin[i], chunk_cmax[i]] (tile order follows the mesh in slabs, so
        // for block-structured numberings chunk i == cell range i): it waits for the U range holding its largest label, and
        // cell range j leaves once the last position chunk that intersects it is done.
        const int NCH = ctx->n_chunk;
        auto range_of = [&](int c) { int j = (int)(((long long)c * NCH) / nC); while ((long long)nC * j / NCH > c) --j; while ((long long)nC * (j + 1) / NCH <= c) ++j; return j; };
        int i_last[sdfibm_context::MAX_CHUNK];
        for (int j = 0; j < NCH; ++j) i_last[j] = -1;
        for (int i = 0; i < NCH; ++i) {
            if (ctx->chunk_cmax[i] < 0) continue;
            for (int j = range_of(ctx->chunk_cmin[i]); j <= range_of(ctx->chunk_cmax[i]); ++j) i_last[j] = i;
        }
        int last_pos_chunk = 0;
        for (int i = 0; i < NCH; ++i) if (ctx->chunk_cmax[i] >= 0) last_pos_chunk = i;
        const int per_slab = std::max(1, NCH / ctx->n_slab);
        CUDA_TRY(ctx->slab_start.ensure(2));
        for (int i = 0; i < NCH; ++i) {
            const long long p0 = (long long)nC * i / NCH, p1 = (long long)nC * (i + 1) / NCH;
            if (i % per_slab == 0) {   // a new slab: classify its positions, evaluate the queue items they add
                const int ie = std::min(NCH, i + per_slab);
                const long long s0 = p0, s1 = (long long)nC * ie / NCH;
                k_snapshot<<<1, 1, 0, st>>>(&ctx->status->heavy_total, &ctx->status->heavy_gen, ctx->slab_start.p);
                I.heavy_start = ctx->slab_start.p;
                if (s1 > s0) launch_classify((int)s0, (int)s1);
                launch_heavy();
                ctx->launches += 3;
            }
            if (p1 > p0) {
                I.c_begin = (int)p0; I.c_end = (int)p1;
                CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_in[range_of(ctx->chunk_cmax[i])], 0));
                k_final<FINAL_CTAS><<<grid_for(p1 - p0, FINAL_NT), FINAL_NT, 0, st>>>(I);
            }
            CUDA_TRY(cudaEventRecord(ctx->ev_fin[i], st));
            for (int j = 0; j < NCH; ++j) {
                if (!(i_last[j] == i || (i_last[j] < 0 && i == last_pos_chunk))) continue;
                const size_t c0 = (size_t)nC * j / NCH, c1 = (size_t)nC * (j + 1) / NCH;
                if (c1 <= c0) continue;
                const size_t n = c1 - c0;
                CUDA_TRY(cudaStreamWaitEvent(ctx->s_out, ctx->ev_fin[i], 0));
                CUDA_TRY(cudaMemcpyAsync(ctx->pipe.As + c0, dAs + c0, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->s_out));
                CUDA_TRY(cudaMemcpyAsync(ctx->pipe.Fs + 3 * c0, dFs + 3 * c0, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->s_out));
                CUDA_TRY(cudaMemcpyAsync(ctx->pipe.Ts + c0, dTs + c0, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->s_out));
                CUDA_TRY(cudaMemcpyAsync(ctx->pipe.Ct + c0, dCt + c0, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->s_out));
            }
        }
        ctx->launches += ctx->n_chunk - 3;
    } else {
        I.c_begin = 0; I.c_end = nC;
        k_final<FINAL_CTAS><<<grid_for(nC, FINAL_NT), FINAL_NT, 0, st>>>(I);
    }
    CUDA_TRY(rec(ctx->ev[4]));
    ctx->launches += 3;
    if (part == 1) {   // the head of a split step: the sums are final once scaled; the tail follows outside the graph
        k_scale_ft<<<std::min(grid_for(6 * (long long)n_solids, 256), 296), 256, 0, st>>>(dFT, 6 * (long long)n_solids, ctx->scal.p);
        ++ctx->launches;
        CUDA_TRY(cudaGetLastError());
        return SDFIBM_OK;
    }
    return enqueue_tail(ctx, n_solids, dFT, replay, capturing, /*scale*/ 1);
}

static void fold_timings(sdfibm_context *ctx) {
    if (!ctx->t_pending) return;
    ctx->t_pending = false;
    for (int k = 0; k < 5; ++k) {
        float x = 0;
        cudaEventElapsedTime(&x, ctx->ev[k], ctx->ev[k + 1]);
        ctx->t_ms[k] = ctx->t_add * ctx->t_ms[k] + x;
    }
    float d = 0;
    cudaEventElapsedTime(&d, ctx->ev[0], ctx->ev[5]);
    ctx->t_ms[5] = ctx->t_add * ctx->t_ms[5] + d;
}

static void drop_graph(sdfibm_context *ctx) {
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    ctx->graph_exec = nullptr;
}

static int run_pipeline(sdfibm_context *ctx, int n_solids, const double *dU, double dt, double rhof, double *dAs,
                        double *dFs, double *dTs, double *dCt, double *dFT, bool replay) {
    cudaStream_t st = ctx->stream;
    const BinGrid &g = ctx->grid;   // the static tile grid
    if (!replay) {
        CUDA_TRY(ctx->solids.ensure(n_solids));
        CUDA_TRY(ctx->bin_off.ensure((size_t)g.n_bins + 1));
        CUDA_TRY(ctx->tile_proven.ensure((size_t)g.n_bins + 1));
        CUDA_TRY(ctx->global_list.ensure(n_solids));
        if (ctx->bin_fixed) CUDA_TRY(ctx->bin_list.ensure((size_t)g.n_bins * BIN_FIXED_CAP));
        else if (ctx->bin_list.n == 0) CUDA_TRY(ctx->bin_list.ensure(std::max<size_t>(1 << 20, 128 * (size_t)n_solids)));
        // one zero-initialised block: [StepStatus | root_count n | pair_counts 3n | bin_count n_bins+1 | bin_cursor n_bins]
        const size_t o_root = (sizeof(StepStatus) + 15) & ~size_t(15);
        const size_t o_pair = o_root + ((sizeof(int) * (size_t)n_solids + 15) & ~size_t(15));
        const size_t o_cnt = o_pair + ((sizeof(unsigned) * 3 * (size_t)n_solids + 15) & ~size_t(15));
        const size_t o_cur = o_cnt + ((sizeof(int) * ((size_t)g.n_bins + 1) + 15) & ~size_t(15));
        const size_t total = o_cur + sizeof(int) * (size_t)g.n_bins;
        if (ctx->zero_block.n != total) { ctx->zero_block.release(); CUDA_TRY(ctx->zero_block.ensure(total)); }
        unsigned char *zb = ctx->zero_block.p;
        ctx->status = reinterpret_cast<StepStatus *>(zb);
        ctx->root_count = reinterpret_cast<int *>(zb + o_root);
        ctx->pair_counts = reinterpret_cast<unsigned *>(zb + o_pair);
        ctx->bin_count = reinterpret_cast<int *>(zb + o_cnt);
        ctx->bin_cursor = reinterpret_cast<int *>(zb + o_cur);
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->bin_count, ctx->bin_off.p, g.n_bins + 1, st);
        CUDA_TRY(ctx->scan_tmp.ensure(tmp_bytes));
    }
    ctx->h_scal[0] = 1.0 / dt;
    ctx->h_scal[1] = rhof;
    if (replay) fold_timings(ctx);   // the first pass of this call: its events are about to be re-recorded
    for (int attempt = 0; attempt < 4; ++attempt) {   // a pass that finds a capacity too small grows it and runs again
        if (!replay) CUDA_TRY(ctx->bin_entries.ensure(ctx->bin_list.n));
        const bool chunked = ctx->pipe.active && !replay && attempt == 0;
        if (ctx->pipe.active && !chunked) {   // a retry / replay pass rewrites the fields: copy them out again at the end
            ctx->pipe.stale = true;
            CUDA_TRY(cudaStreamSynchronize(ctx->s_in));
            CUDA_TRY(cudaStreamSynchronize(ctx->s_out));
        }
        const bool use_graph = ctx->use_graph && !replay && !ctx->pipe.active;
        const bool reduced_here = ctx->reduce_out && attempt == 0 && !replay;
        // split step (multi-GPU): the graph ends behind k_final + the rhof scaling; the all-reduce of the sums then runs on the
        // communication stream WHILE the certificate pass and the status totals follow on the context stream
        const bool split = use_graph && reduced_here && ctx->comm_split && ctx->s_comm;
        const auto q0 = std::chrono::steady_clock::now();
        if (use_graph) {
            const uint64_t key[20] = {(uint64_t)n_solids, (uint64_t)dU, (uint64_t)dAs, (uint64_t)dFs, (uint64_t)dTs, (uint64_t)dCt, (uint64_t)dFT,
                                      (uint64_t)(ctx->ext_solids ? ctx->ext_solids : (ctx->gathered_now ? ctx->solids_gathered.p : ctx->solids_in.p)), (uint64_t)ctx->solids.p, (uint64_t)ctx->heavy.p, (uint64_t)ctx->heavy.n,
                                      (uint64_t)ctx->bin_list.p, (uint64_t)ctx->bin_list.n, (uint64_t)ctx->bin_entries.p, (uint64_t)ctx->zero_block.p,
                                      (uint64_t)ctx->shapes.p ^ ((uint64_t)ctx->sdf_ops.p << 2), (uint64_t)ctx->scan_tmp.p ^ ((uint64_t)ctx->tile_proven.p << 1), (uint64_t)ctx->global_list.p, (uint64_t)ctx->shapes_refinable,
                                      (uint64_t)ctx->n_global_hint ^ ((uint64_t)ctx->bin_fixed << 8) ^ ((uint64_t)split << 9)};
            if (!ctx->graph_exec || memcmp(key, ctx->graph_key, sizeof(key)) != 0) {
                drop_graph(ctx);
                cudaGraph_t graph = nullptr;
                CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                const int64_t keep_launches = ctx->launches;
                const int rc = enqueue_pipeline(ctx, n_solids, dU, dAs, dFs, dTs, dCt, dFT, false, false, true, split ? 1 : 0);
                ctx->graph_launches = ctx->launches - keep_launches;
                ctx->launches = keep_launches;
                const cudaError_t e = cudaStreamEndCapture(st, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (e != cudaSuccess) return fail(SDFIBM_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
                const cudaError_t e2 = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
                cudaGraphDestroy(graph);
                if (e2 != cudaSuccess) { ctx->graph_exec = nullptr; return fail(SDFIBM_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2)); }
                memcpy(ctx->graph_key, key, sizeof(key));
            }
            CUDA_TRY(cudaGraphLaunch(ctx->graph_exec, st));
            ctx->launches += ctx->graph_launches;
            if (split) {
                NcclApi *a = nccl_api();
                CUDA_TRY(cudaEventRecord(ctx->ev_ft, st));                        // the scaled sums are final
                CUDA_TRY(cudaStreamWaitEvent(ctx->s_comm, ctx->ev_ft, 0));
                CUDA_TRY(cudaEventRecord(ctx->ev_comm[0], ctx->s_comm));
                NCCL_TRY(a->AllReduce(dFT, ctx->reduce_out, 6 * (size_t)n_solids, ncclDouble, ncclSum, ctx->comm, ctx->s_comm));
                CUDA_TRY(cudaEventRecord(ctx->ev_comm[1], ctx->s_comm));
                const int rc = enqueue_tail(ctx, n_solids, dFT, false, false, /*scale*/ 0);   // certificate + totals, alongside the all-reduce
                if (rc) return rc;
                // the ranks' "again" flags need the certificate's result: a second, 8-byte all-reduce behind the tail
                k_retry_flag<<<1, 1, 0, st>>>(ctx->status, (long long)ctx->heavy.n, ctx->n_global_hint, ctx->retry_flag.p);
                NCCL_TRY(a->AllReduce(ctx->retry_flag.p, ctx->retry_flag.p + 1, 1, ncclDouble, ncclSum, ctx->comm, st));
                CUDA_TRY(cudaMemcpyAsync(ctx->h_retry_sum, ctx->retry_flag.p + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_comm[1], 0));             // the host's one synchronisation covers both streams
            }
        } else {
            const int rc = enqueue_pipeline(ctx, n_solids, dU, dAs, dFs, dTs, dCt, dFT, replay, chunked, false);
            if (rc) return rc;
        }
        if (reduced_here && !split) {
            const int rc = enqueue_comm_reduce(ctx, dFT, ctx->reduce_out, n_solids, true);
            if (rc) return rc;
        }
        const auto w0 = std::chrono::steady_clock::now();
        CUDA_TRY(cudaStreamSynchronize(st));
        if (reduced_here) { float x = 0; cudaEventElapsedTime(&x, ctx->ev_comm[0], ctx->ev_comm[1]); ctx->t_comm_ms = x; }
        ctx->t_host_us[2] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - w0).count();
        ctx->t_host_us[1] = std::chrono::duration<double, std::micro>(w0 - q0).count();
        // the event times of this pass are read when somebody asks for them (sdfibm_last_timings) or before a replay pass re-records
        // the events: six cudaEventElapsedTime calls are ~10 us of host time per step otherwise
        ctx->t_add = replay ? 1.0 : 0.0;   // a replay pass adds to the first pass of the same call
        ctx->t_pending = true;
        {
            ctx->last = *ctx->h_status;
            if (replay) ctx->last.n_flagged = (int)ctx->flagged_last;
        }
        bool again = false;
        if (ctx->last.n_global > 0 && !ctx->n_global_hint && !replay) {
            // the caller of sdfibm_interact_device_solids promised "no plane / tilted 2-D solid" and was wrong: the classify variant
            // that was run never looked at the global list.  Run again with the variant that does.
            ctx->n_global_hint = 1;
            again = true;
        }
        if (ctx->last.slot_overflow && !replay && ctx->K < 64) {
            // more solids touch one cell than there are slot records (dense packs against a wall plane): widen and run again
            const int K = std::min(64, std::max(ctx->last.slot_need, ctx->K + 1));
            drop_graph(ctx);
            CUDA_TRY(ctx->slots.ensure((size_t)ctx->dm.n_cells * K));
            ctx->K = K;
            again = true;
        }
        if (ctx->last.heavy_total + ctx->last.heavy_gen > (unsigned long long)ctx->heavy.n) {
            const unsigned long long need = ctx->last.heavy_total + ctx->last.heavy_gen;
            if (need + need / 4 + 1024 >= (1ull << 28)) return fail(SDFIBM_ERR_CAPACITY, "exact-evaluation queue would exceed 2^28 items (slot records hold the queue index in 28 bits): split the mesh");
            const size_t cap = (size_t)(need + need / 4 + 1024);
            CUDA_TRY(ctx->heavy.ensure(cap));
            CUDA_TRY(ctx->heavy_res.ensure(cap));
            again = true;
        }
        if (ctx->last.bin_overflow && !replay) {
            if (ctx->bin_fixed) { ctx->bin_fixed = false; drop_graph(ctx); }   // a tile with more candidates than slots: scan + fill from now on
            CUDA_TRY(ctx->bin_list.ensure((size_t)ctx->last.bin_total + (size_t)ctx->last.bin_total / 4 + 1024));
            again = true;
        }
        if (!again || attempt == 3) break;
    }
    if (ctx->last.bin_overflow) return fail(SDFIBM_ERR_CAPACITY, "solid bin list overflow");
    if (ctx->last.bad_cell) return fail(SDFIBM_ERR_UNSUPPORTED, "cell with more than 32 vertices");
    if (ctx->last.bad_shape) return fail(SDFIBM_ERR_ARG, "solid refers to an unknown shape index");
    if (ctx->last.slot_overflow)
        return fail(SDFIBM_ERR_CAPACITY, "more than 64 solids touch one cell (slot records)");
    if (ctx->last.heavy_total + ctx->last.heavy_gen > (unsigned long long)ctx->heavy.n) return fail(SDFIBM_ERR_CAPACITY, "exact-evaluation queue overflow");
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
    if (!ctx || n_solids < 0 || (n_solids > 0 && (!solids || !force_torque)) || !U || !As || !Fs || !Ts || !Ct)
        return fail(SDFIBM_ERR_ARG, "sdfibm_interact: null argument");
    if (!ctx->has_mesh || (n_solids > 0 && ctx->h_shapes.empty())) return fail(SDFIBM_ERR_STATE, "sdfibm_interact: set mesh and shapes first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_fields(ctx, std::max(n_solids, 1));
    if (rc) return rc;
    const size_t nC = ctx->dm.n_cells;
    cudaStream_t st = ctx->stream;
    if (n_solids == 0) {
        rc = sdfibm_interact_device(ctx, nullptr, 0, ctx->dU.p, dt, rhof, ctx->dAs.p, ctx->dFs.p, ctx->dTs.p, ctx->dCt.p, nullptr);
        if (rc) return rc;
        memset(As, 0, sizeof(double) * nC); memset(Fs, 0, sizeof(double) * 3 * nC);
        memset(Ts, 0, sizeof(double) * nC); memset(Ct, 0, sizeof(double) * nC);
        return SDFIBM_OK;
    }
    // the copy engines are FIFO across streams: the small solid upload the kernels depend on goes first
    rc = stage_solids(ctx, solids, n_solids, true);
    if (rc) return rc;
    // U is not needed before k_final: its chunks stream in on s_in while binning / k_classify / k_heavy run; each chunk of the
    // fields leaves on s_out as soon as its k_final launch is done (PCIe is full duplex: copy-in and copy-out overlap too).
    ctx->pipe.active = true; ctx->pipe.stale = false;
    ctx->pipe.As = As; ctx->pipe.Fs = Fs; ctx->pipe.Ts = Ts; ctx->pipe.Ct = Ct; ctx->pipe.U = U;
    rc = sdfibm_interact_device(ctx, solids, n_solids, ctx->dU.p, dt, rhof, ctx->dAs.p, ctx->dFs.p, ctx->dTs.p, ctx->dCt.p, ctx->dFT.p);
    ctx->pipe.active = false;
    if (rc) { cudaStreamSynchronize(ctx->s_in); cudaStreamSynchronize(ctx->s_out); return rc; }
    CUDA_TRY(cudaMemcpyAsync(force_torque, ctx->dFT.p, sizeof(double) * 6 * n_solids, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_out));
    if (ctx->pipe.stale) {
        CUDA_TRY(cudaMemcpyAsync(As, ctx->dAs.p, sizeof(double) * nC, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(Fs, ctx->dFs.p, sizeof(double) * 3 * nC, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(Ts, ctx->dTs.p, sizeof(double) * nC, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(Ct, ctx->dCt.p, sizeof(double) * nC, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return SDFIBM_OK;
}

int sdfibm_interact_device_solids(sdfibm_context *ctx, const sdfibm_solid_t *d_solids, int n_solids, int may_be_global, const double *dU,
                                  double dt, double rhof, double *dAs, double *dFs, double *dTs, double *dCt, double *dFT) {
    if (!ctx || (n_solids > 0 && !d_solids)) return fail(SDFIBM_ERR_ARG, "sdfibm_interact_device_solids: null argument");
    ctx->ext_solids = n_solids > 0 ? d_solids : nullptr;
    ctx->n_global_hint = (may_be_global && ctx->shapes_may_be_global) ? 1 : 0;
    const int rc = sdfibm_interact_device(ctx, nullptr, n_solids, dU, dt, rhof, dAs, dFs, dTs, dCt, dFT);
    ctx->ext_solids = nullptr;
    return rc;
}

// ---- the step either side of interact, on the device (src/main.cpp:70-77) ----
int sdfibm_apply_forcing_device(sdfibm_context *ctx, double *dU, double *dT, double dt) {
    if (!ctx || (!dU && !dT)) return fail(SDFIBM_ERR_ARG, "sdfibm_apply_forcing_device: null argument");
    if (!ctx->has_mesh || !ctx->last_Fs) return fail(SDFIBM_ERR_STATE, "sdfibm_apply_forcing_device: no interact has run on this context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int nC = ctx->dm.n_cells;
    k_apply_forcing<<<grid_for(nC, 256), 256, 0, ctx->stream>>>(ctx->n_item.p, ctx->orig.p, nC, ctx->last_As, ctx->last_Fs, ctx->last_Ts, dt, dU, dT);
    CUDA_TRY(cudaGetLastError());
    return SDFIBM_OK;   // stream-ordered: the next call on this context (or sdfibm_synchronize) sees the result
}

int sdfibm_download(sdfibm_context *ctx, void *host_dst, const void *device_src, size_t bytes) {
    if (!ctx || (bytes && (!host_dst || !device_src))) return fail(SDFIBM_ERR_ARG, "sdfibm_download: null argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (bytes) CUDA_TRY(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SDFIBM_OK;
}

int sdfibm_touched_cells(sdfibm_context *ctx, int64_t capacity, int64_t *n_touched, int32_t *cells, double *As, double *Fs, double *Ts, double *Ct) {
    if (!ctx || !n_touched) return fail(SDFIBM_ERR_ARG, "sdfibm_touched_cells: null argument");
    if (!ctx->has_mesh || !ctx->last_Fs) return fail(SDFIBM_ERR_STATE, "sdfibm_touched_cells: no interact has run on this context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int nC = ctx->dm.n_cells;
    CUDA_TRY(ctx->t_flag.ensure((size_t)nC + 1));
    CUDA_TRY(ctx->t_off.ensure((size_t)nC + 1));
    k_touched_flags<<<grid_for((long long)nC + 1, 256), 256, 0, st>>>(ctx->n_item.p, nC, ctx->t_flag.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, ctx->t_flag.p, ctx->t_off.p, nC + 1, st);
    CUDA_TRY(ctx->scan_tmp.ensure(tb));
    cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tb, ctx->t_flag.p, ctx->t_off.p, nC + 1, st);
    int total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, ctx->t_off.p + nC, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *n_touched = total;
    if (!cells) return SDFIBM_OK;   // count only
    if (!As || !Fs || !Ts || !Ct) return fail(SDFIBM_ERR_ARG, "sdfibm_touched_cells: null output array");
    if (total > capacity) return fail(SDFIBM_ERR_CAPACITY, "sdfibm_touched_cells: capacity too small");
    if (total == 0) return SDFIBM_OK;
    const size_t n = (size_t)total;
    CUDA_TRY(ctx->t_cells.ensure(n));
    CUDA_TRY(ctx->t_vals.ensure(6 * n));
    double *oAs = ctx->t_vals.p, *oFs = oAs + n, *oTs = oFs + 3 * n, *oCt = oTs + n;
    k_touched_gather<<<grid_for(nC, 256), 256, 0, st>>>(ctx->n_item.p, ctx->orig.p, ctx->t_off.p, nC, (long long)n, ctx->last_As, ctx->last_Fs,
                                                       ctx->last_Ts, ctx->last_Ct, ctx->t_cells.p, oAs, oFs, oTs, oCt);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(cells, ctx->t_cells.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(As, oAs, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(Fs, oFs, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(Ts, oTs, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(Ct, oCt, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SDFIBM_OK;
}

// ---- cross-rank exchange (NCCL over NVLink) ----
int sdfibm_comm_unique_id(void *id) {
    if (!id) return fail(SDFIBM_ERR_ARG, "sdfibm_comm_unique_id: null argument");
    NcclApi *a = nccl_api();
    if (!a->handle) return fail(SDFIBM_ERR_UNSUPPORTED, a->error);
    static_assert(sizeof(ncclUniqueId) == SDFIBM_COMM_ID_BYTES, "ncclUniqueId size");
    NCCL_TRY(a->GetUniqueId(reinterpret_cast<ncclUniqueId *>(id)));
    return SDFIBM_OK;
}

int sdfibm_comm_init(sdfibm_context *ctx, const void *id, int rank, int n_ranks) {
    if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(SDFIBM_ERR_ARG, "sdfibm_comm_init: bad argument");
    if (ctx->comm) return fail(SDFIBM_ERR_STATE, "sdfibm_comm_init: the context already has a communicator");
    NcclApi *a = nccl_api();
    if (!a->handle) return fail(SDFIBM_ERR_UNSUPPORTED, a->error);
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NCCL_TRY(a->CommInitRank(&ctx->comm, n_ranks, uid, rank));
    ctx->comm_rank = rank;
    ctx->comm_n = n_ranks;
    ctx->comm_auto_reduce = true;
    ctx->comm_gather_solids = true;
    drop_graph(ctx);
    return SDFIBM_OK;
}

int sdfibm_comm_options(sdfibm_context *ctx, int auto_reduce, int gather_solids) {
    if (!ctx) return fail(SDFIBM_ERR_ARG, "sdfibm_comm_options: null context");
    ctx->comm_auto_reduce = auto_reduce != 0;
    ctx->comm_gather_solids = gather_solids != 0;
    drop_graph(ctx);
    return SDFIBM_OK;
}

int sdfibm_comm_destroy(sdfibm_context *ctx) {
    if (!ctx) return fail(SDFIBM_ERR_ARG, "sdfibm_comm_destroy: null context");
    if (ctx->comm) {
        CUDA_TRY(cudaSetDevice(ctx->device));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->s_comm) CUDA_TRY(cudaStreamSynchronize(ctx->s_comm));
        NCCL_TRY(nccl_api()->CommDestroy(ctx->comm));
    }
    ctx->comm = nullptr;
    ctx->comm_n = 1;
    ctx->comm_rank = 0;
    return SDFIBM_OK;
}

int sdfibm_allreduce_force_torque(sdfibm_context *ctx, double *d_force_torque, int n_solids) {
    if (!ctx || !d_force_torque || n_solids < 0) return fail(SDFIBM_ERR_ARG, "sdfibm_allreduce_force_torque: bad argument");
    if (!ctx->comm) return fail(SDFIBM_ERR_STATE, "sdfibm_allreduce_force_torque: call sdfibm_comm_init first");
    if (n_solids == 0) return SDFIBM_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return enqueue_comm_reduce(ctx, d_force_torque, d_force_torque, n_solids, false);
}

int sdfibm_comm_last_ms(sdfibm_context *ctx, double *ms) {
    if (!ctx || !ms) return fail(SDFIBM_ERR_ARG, "null argument");
    *ms = ctx->t_comm_ms;
    return SDFIBM_OK;
}

int sdfibm_fix_internal_device(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, double *dU, const double *dCt) {
    if (!ctx || n_solids < 0 || (n_solids > 0 && !solids) || !dU) return fail(SDFIBM_ERR_ARG, "sdfibm_fix_internal: null argument");
    if (n_solids == 0) return SDFIBM_OK;   // Ct >= 4 nowhere
    if (!ctx->has_mesh) return fail(SDFIBM_ERR_STATE, "sdfibm_fix_internal: set mesh first");
    if (!dCt) dCt = ctx->last_Ct;
    if (!dCt) return fail(SDFIBM_ERR_STATE, "sdfibm_fix_internal: no Ct available (call interact first)");
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = stage_solids(ctx, solids, n_solids);
    if (rc) return rc;
    for (int i = 0; i < 2; ++i) if (!ctx->ev_aux[i]) CUDA_TRY(cudaEventCreate(&ctx->ev_aux[i]));
    CUDA_TRY(cudaEventRecord(ctx->ev_aux[0], ctx->stream));
    k_fix_internal<<<grid_for(((long long)ctx->dm.n_cells + 1) / 2, 256), 256, 0, ctx->stream>>>(ctx->cc_orig.p, ctx->solids_in.p, n_solids, dCt, dU, 0, ctx->dm.n_cells);
    CUDA_TRY(cudaEventRecord(ctx->ev_aux[1], ctx->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    { float x = 0; cudaEventElapsedTime(&x, ctx->ev_aux[0], ctx->ev_aux[1]); ctx->t_fix_ms = x; }
    return SDFIBM_OK;
}

int sdfibm_fix_internal(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, double *U) {
    if (!ctx || !U || n_solids < 0 || (n_solids > 0 && !solids)) return fail(SDFIBM_ERR_ARG, "sdfibm_fix_internal: null argument");
    if (n_solids == 0) return SDFIBM_OK;
    if (!ctx->has_mesh || !ctx->dCt.p || ctx->last_Ct != ctx->dCt.p)
        return fail(SDFIBM_ERR_STATE, "sdfibm_fix_internal: call sdfibm_interact (host buffers) first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t nC = ctx->dm.n_cells;
    cudaStream_t st = ctx->stream;
    int rc = stage_solids(ctx, solids, n_solids);   // first on the copy engine, ahead of the U chunks
    if (rc) return rc;
    // U in / kernel / U out per cell chunk: the two copy directions overlap (full-duplex PCIe)
    for (int i = 0; i < ctx->n_chunk; ++i) {
        const size_t c0 = nC * i / ctx->n_chunk, c1 = nC * (i + 1) / ctx->n_chunk;
        if (c1 > c0) CUDA_TRY(cudaMemcpyAsync(ctx->dU.p + 3 * c0, U + 3 * c0, sizeof(double) * 3 * (c1 - c0), cudaMemcpyHostToDevice, ctx->s_in));
        CUDA_TRY(cudaEventRecord(ctx->ev_in[i], ctx->s_in));
    }
    for (int i = 0; i < ctx->n_chunk; ++i) {
        const size_t c0 = nC * i / ctx->n_chunk, c1 = nC * (i + 1) / ctx->n_chunk;
        if (c1 <= c0) continue;
        CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_in[i], 0));
        k_fix_internal<<<grid_for(((long long)(c1 - c0) + 1) / 2, 256), 256, 0, st>>>(ctx->cc_orig.p, ctx->solids_in.p, n_solids, ctx->dCt.p, ctx->dU.p, (int)c0, (int)c1);
        CUDA_TRY(cudaEventRecord(ctx->ev_fin[i], st));
        CUDA_TRY(cudaStreamWaitEvent(ctx->s_out, ctx->ev_fin[i], 0));
        CUDA_TRY(cudaMemcpyAsync(U + 3 * c0, ctx->dU.p + 3 * c0, sizeof(double) * 3 * (c1 - c0), cudaMemcpyDeviceToHost, ctx->s_out));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->s_out));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SDFIBM_OK;
}

__global__ void k_fill_unit_x(double *U, long long n_cells) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    U[3 * c] = 1.0; U[3 * c + 1] = 0.0; U[3 * c + 2] = 0.0;
}

int sdfibm_mean_field_sums(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, const double *field, double *sum_alpha_v_field,
                           double *sum_alpha_v) {
    if (!ctx || !solids || n_solids <= 0 || !field || !sum_alpha_v_field || !sum_alpha_v) return fail(SDFIBM_ERR_ARG, "sdfibm_mean_field: null argument");
    if (!ctx->has_mesh || ctx->h_shapes.empty()) return fail(SDFIBM_ERR_STATE, "sdfibm_mean_field: set mesh and shapes first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t nC = ctx->dm.n_cells;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx->sU.ensure(3 * nC));
    CUDA_TRY(ctx->sOut.ensure(6 * nC));
    CUDA_TRY(ctx->sFT.ensure(12 * (size_t)n_solids));
    // solids at rest, dt = 1, rhof = 1: interact's per-solid "force" is sum(alpha V U) (solidcloud.cpp:414-416)
    std::vector<sdfibm_solid_t> rest(solids, solids + n_solids);
    for (auto &r : rest)
        for (int d = 0; d < 3; ++d) r.vel[d] = r.omega[d] = 0.0;
    double *tAs = ctx->sOut.p, *tFs = tAs + nC, *tTs = tFs + 3 * nC, *tCt = tTs + nC;
    const double *keep_Ct = ctx->last_Ct;
    const int keep_n = ctx->n_solids_last;
    CUDA_TRY(cudaMemcpyAsync(ctx->sU.p, field, sizeof(double) * 3 * nC, cudaMemcpyHostToDevice, st));
    const bool keep_reduce = ctx->comm_auto_reduce, keep_gather = ctx->comm_gather_solids;
    ctx->comm_auto_reduce = false; ctx->comm_gather_solids = false;   // this rank's sums: the caller reduces numerator and denominator
    int rc = sdfibm_interact_device(ctx, rest.data(), n_solids, ctx->sU.p, 1.0, 1.0, tAs, tFs, tTs, tCt, ctx->sFT.p);
    if (!rc) {
        k_fill_unit_x<<<grid_for((long long)nC, 256), 256, 0, st>>>(ctx->sU.p, (long long)nC);
        rc = sdfibm_interact_device(ctx, rest.data(), n_solids, ctx->sU.p, 1.0, 1.0, tAs, tFs, tTs, tCt, ctx->sFT.p + 6 * (size_t)n_solids);
    }
    ctx->comm_auto_reduce = keep_reduce; ctx->comm_gather_solids = keep_gather;
    ctx->last_Ct = keep_Ct;   // the sampler leaves the coupling state of the last interact alone (its candidate lists are replaced)
    ctx->n_solids_last = keep_n;
    if (rc) return rc;
    std::vector<double> ft(12 * (size_t)n_solids);
    CUDA_TRY(cudaMemcpyAsync(ft.data(), ctx->sFT.p, sizeof(double) * ft.size(), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int s = 0; s < n_solids; ++s) {
        for (int d = 0; d < 3; ++d) sum_alpha_v_field[3 * s + d] = ft[6 * (size_t)s + d];
        sum_alpha_v[s] = ft[6 * ((size_t)n_solids + s)];
    }
    return SDFIBM_OK;
}

// tool_vof (tool_vof/solidcloud.cpp:116-173): the solid volume fraction of every cell — interact's As kernel on scratch outputs
int sdfibm_volume_fraction(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, double *alpha, double *total_volume) {
    if (!ctx || n_solids < 0 || (n_solids > 0 && !solids) || !alpha) return fail(SDFIBM_ERR_ARG, "sdfibm_volume_fraction: null argument");
    if (!ctx->has_mesh || (n_solids > 0 && ctx->h_shapes.empty())) return fail(SDFIBM_ERR_STATE, "sdfibm_volume_fraction: set mesh and shapes first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t nC = ctx->dm.n_cells;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx->sU.ensure(3 * nC));
    CUDA_TRY(ctx->sOut.ensure(6 * nC));
    CUDA_TRY(ctx->sFT.ensure(12 * (size_t)std::max(n_solids, 1)));
    // solids at rest in a fluid at rest, dt = 1: only As is of interest (the per-solid += and the clamp at 1 of
    // tool_vof/solidcloud.cpp:131-133 are interact's own accumulation order and checkAlpha)
    std::vector<sdfibm_solid_t> rest(solids, solids + n_solids);
    for (auto &r : rest)
        for (int d = 0; d < 3; ++d) r.vel[d] = r.omega[d] = 0.0;
    double *tAs = ctx->sOut.p, *tFs = tAs + nC, *tTs = tFs + 3 * nC, *tCt = tTs + nC;
    const double *keep_Ct = ctx->last_Ct, *keep_As = ctx->last_As, *keep_Fs = ctx->last_Fs, *keep_Ts = ctx->last_Ts;
    const int keep_n = ctx->n_solids_last;
    const bool keep_reduce = ctx->comm_auto_reduce, keep_gather = ctx->comm_gather_solids;
    ctx->comm_auto_reduce = false; ctx->comm_gather_solids = false;
    CUDA_TRY(cudaMemsetAsync(ctx->sU.p, 0, sizeof(double) * 3 * nC, st));
    const int rc = sdfibm_interact_device(ctx, rest.data(), n_solids, ctx->sU.p, 1.0, 1.0, tAs, tFs, tTs, tCt, ctx->sFT.p);
    ctx->comm_auto_reduce = keep_reduce; ctx->comm_gather_solids = keep_gather;
    ctx->last_Ct = keep_Ct; ctx->last_As = keep_As; ctx->last_Fs = keep_Fs; ctx->last_Ts = keep_Ts;
    ctx->n_solids_last = keep_n;
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(alpha, tAs, sizeof(double) * nC, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (total_volume) {   // gSum(tmp * V) (:169)
        std::vector<double> V(nC);
        std::vector<int> orig(nC);
        CUDA_TRY(cudaMemcpy(V.data(), ctx->V.p, sizeof(double) * nC, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(orig.data(), ctx->orig.p, sizeof(int) * nC, cudaMemcpyDeviceToHost));
        double sum = 0.0;
        std::vector<double> Vc(nC);
        for (size_t i = 0; i < nC; ++i) Vc[orig[i]] = V[i];
        for (size_t c = 0; c < nC; ++c) sum += alpha[c] * Vc[c];
        *total_volume = sum;
    }
    return SDFIBM_OK;
}

int sdfibm_mean_field(sdfibm_context *ctx, const sdfibm_solid_t *solids, int n_solids, const double *field, double *mean,
                      double *sum_alpha_v) {
    if (!mean || n_solids <= 0) return fail(SDFIBM_ERR_ARG, "sdfibm_mean_field: null argument");
    std::vector<double> den((size_t)n_solids);
    const int rc = sdfibm_mean_field_sums(ctx, solids, n_solids, field, mean, den.data());
    if (rc) return rc;
    for (int s = 0; s < n_solids; ++s) {
        for (int d = 0; d < 3; ++d) mean[3 * s + d] /= den[s];   // single rank: 0/0 = NaN for a solid that touches no cell, as in the reference
        if (sum_alpha_v) sum_alpha_v[s] = den[s];
    }
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
    stats[3] = (int64_t)(ctx->last.heavy_total + ctx->last.heavy_gen);
    return SDFIBM_OK;
}

int sdfibm_last_timings(sdfibm_context *ctx, double ms[6]) {
    if (!ctx || !ms) return fail(SDFIBM_ERR_ARG, "null argument");
    fold_timings(ctx);
    for (int k = 0; k < 6; ++k) ms[k] = ctx->t_ms[k];
    return SDFIBM_OK;
}

int sdfibm_last_aux_timings(sdfibm_context *ctx, double ms[2]) {
    if (!ctx || !ms) return fail(SDFIBM_ERR_ARG, "sdfibm_last_aux_timings: null argument");
    ms[0] = ctx->t_fix_ms;
    ms[1] = ctx->t_col_ms;
    return SDFIBM_OK;
}

int sdfibm_last_host_timings(sdfibm_context *ctx, double us[4]) {
    if (!ctx || !us) return fail(SDFIBM_ERR_ARG, "null argument");
    for (int k = 0; k < 4; ++k) us[k] = ctx->t_host_us[k];
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
    CUDA_TRY(cudaMemcpyAsync(h_cnt.data(), ctx->pair_counts, sizeof(unsigned) * 3 * nS, cudaMemcpyDeviceToHost, st));
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
    k_list_count<<<grid_for(nC, 256), 256, 0, st>>>(ctx->n_item.p, ctx->slots.p, excl, ctx->K, nC, cnt.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, off.p, nC + 1, st);
    CUDA_TRY(tmp.ensure(tb));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.p, off.p, nC + 1, st);
    k_list_emit<<<grid_for(nC, 256), 256, 0, st>>>(ctx->n_item.p, ctx->slots.p, excl, ctx->K, nC, off.p, ctx->orig.p, keys.p, vals.p);
    size_t sb = 0, sb2 = 0;
    // stable sort by the caller's cell label, then stable sort by (solid, type)
    cub::DeviceRadixSort::SortPairs(nullptr, sb, (unsigned *)vals.p, (unsigned *)vals2.p, keys.p, keys2.p, (int)total, 0, 32, st);
    cub::DeviceRadixSort::SortPairs(nullptr, sb2, keys2.p, keys.p, vals2.p, vals.p, (int)total, 0, 32, st);
    CUDA_TRY(tmp.ensure(std::max(sb, sb2)));
    cub::DeviceRadixSort::SortPairs(tmp.p, sb, (unsigned *)vals.p, (unsigned *)vals2.p, keys.p, keys2.p, (int)total, 0, 32, st);
    cub::DeviceRadixSort::SortPairs(tmp.p, sb2, keys2.p, keys.p, vals2.p, vals.p, (int)total, 0, 32, st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(cells, vals.p, sizeof(int) * total, cudaMemcpyDeviceToHost, st));
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
    DevBuf<int> &ids = ctx->col_ids, &sids = ctx->col_sids, &pcnt = ctx->col_pcnt, &poff = ctx->col_poff, &dpairs = ctx->col_pairs;
    DevBuf<unsigned> &keys = ctx->col_keys, &skeys = ctx->col_skeys;
    DevBuf<unsigned char> &tmp = ctx->col_tmp;
    DevBuf<double> &dft = ctx->col_ft;
    for (int i = 0; i < 2; ++i) if (!ctx->ev_aux[i]) CUDA_TRY(cudaEventCreate(&ctx->ev_aux[i]));
    size_t sb = 0, tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, keys.p, skeys.p, ids.p, sids.p, n, 0, 32, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, pcnt.p, poff.p, n + 1, st);
    CUDA_TRY(tmp.ensure(std::max(sb, tb)));
    CUDA_TRY(cudaEventRecord(ctx->ev_aux[0], st));
    CUDA_TRY(keys.ensure(n)); CUDA_TRY(ids.ensure(n)); CUDA_TRY(skeys.ensure(n)); CUDA_TRY(sids.ensure(n));
    CUDA_TRY(pcnt.ensure((size_t)n + 1)); CUDA_TRY(poff.ensure((size_t)n + 1));
    k_col_keys<<<grid_for(n, 256), 256, 0, st>>>(ctx->solids_in.p, n, g, keys.p, ids.p);
    cub::DeviceRadixSort::SortPairs(tmp.p, sb, keys.p, skeys.p, ids.p, sids.p, n, 0, 32, st); // stable; keys are signed but compared as bits:
    CUDA_TRY(cudaMemsetAsync(pcnt.p, 0, sizeof(int) * ((size_t)n + 1), st));
    k_col_pairs<<<grid_for(n, 128), 128, 0, st>>>(skeys.p, sids.p, n, g, pcnt.p, nullptr, nullptr, 0, 0);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, pcnt.p, poff.p, n + 1, st);
    int *h_total = reinterpret_cast<int *>(ctx->h_scal + 2);   // page-locked scratch next to the per-step scalars
    CUDA_TRY(cudaMemcpyAsync(h_total, poff.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int total = *h_total;
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
    CUDA_TRY(cudaEventRecord(ctx->ev_aux[1], st));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    { float x = 0; cudaEventElapsedTime(&x, ctx->ev_aux[0], ctx->ev_aux[1]); ctx->t_col_ms = x; }
    return SDFIBM_OK;
}

} // extern "C"
