// interact_kernels.cuh — the sm_100a kernels of SolidCloud::interact / fixInternal.
// Included by sdfibm_cuda.cu after the device records (DevMesh, DevSolid, DevShape, BinGrid, StepStatus).
//
//   k_classify      thread per cell   candidate solids (ascending id) -> per-cell slot record; pairs whose vertices must be
//                                     evaluated exactly are appended to a dense global queue (block-aggregated: one atomic per CTA)
//   k_heavy         lane per item     exact vertex predicates + SDF, cell type, apex/pyramid volume (dense, no barriers)
//   k_final         thread per cell   As/Fs/Ts/Ct in the reference's += order, every cell written exactly once in full,
//                                     coalesced sectors; per-solid force/torque warp-aggregated before the atomics
//   k_connectivity  thread per cell   certificate that each solid's cell set is one face-connected component
//   k_replay_*                        exact flood-fill component selection for solids that fail it (rare)
//   k_fix_internal  thread per cell   SolidCloud::fixInternal
//   k_list_*                          candidate-list extraction for parity (off the timed path)
//
// Slot records: slots[j * n_cells + c], j < n_item[c] <= K.  After k_classify a slot is (solid << 3) | ALL_INSIDE for a
// pair the pre-classification proved inside, or (queue index << 3) | 4 for a queued pair.  k_final rewrites every slot
// as (solid << 3) | (queued ? 4 : 0) | type with the final CELL_TYPE (1,2,3) or 0 (no vertex inside: not a member).
#pragma once

#define TPB 128
#ifndef CLS4_NT
#define CLS4_NT 64      // k_classify4: threads per CTA (x 4 positions)
#endif
#ifndef CLS4_PREFETCH
#define CLS4_PREFETCH 0   // requesting the next candidate record while the current one is tested: measured slower (spills at 64 registers)
#endif
#ifndef CLS4_MINB
#define CLS4_MINB 16
#endif
#define SLOT_HEAVY 4
#define BIN_FIXED_CAP 8   // list slots per tile in the fixed-capacity binning mode (a tile with more candidates switches the context to scan + fill)

// ------------------------------------------------------------------------------------------------
// geometry: apex / pyramid volume fraction (reference src/geometrictools.cpp:13-116)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double line_fraction(double a, double b) {          // :13-23
    // branch-free: ONE division per call for the whole warp.  in = the phi <= 0 end, out = the phi > 0 end; the quotient is
    // the reference's -phi_in / (phi_out - phi_in) and is discarded when both ends lie on the same side.
    const bool ap = a > 0, bp = b > 0;
    const double in = ap ? b : a, out = ap ? a : b;
    const double q = -in / (out - in);
    return (ap && bp) ? 0.0 : ((ap || bp) ? q : 1.0);
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

// calcCellVolume (:47-72) for cell c with cell-local vertex coordinates and phi already evaluated (general polyhedra).
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
            eps_f = area / __ldg(m.magSf + f);
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
// parameters shared by the classify / heavy / final kernels
// ------------------------------------------------------------------------------------------------
// one candidate of a tile: what the fp32 sphere-type pre-classification needs, inline (32 bytes, two 16-byte loads)
struct __align__(16) BinEntry {
    float x, y, z;     // solid centre relative to the mesh origin
    float r_out;       // certified outer radius + fp32 slack, rounded up
    float r_in;        // certified inner radius - fp32 slack, rounded down
    int s;             // solid id
    int kind;
    int refine;        // DevSolid::refine of the solid: 0 = no fp32 corner refinement for this shape
};

struct InteractParams {
    DevMesh m;
    const DevSolid *solids;
    const DevShape *shapes;
    const sdfibm_sdf_op_t *ops;   // op table of the SDFIBM_SHAPE_PROGRAM records (null: none)
    int n_solids;
    BinGrid grid;
    int part_shapes;             // exact-box meshes: queued pairs of un-rotated spheres go to the FRONT queue, all others to the BACK queue
                                 // (item i at heavy_cap - 1 - i), so that the warps of k_heavy_box are uniform in their corner path
    int bin_fixed;               // 1: tile t owns the slots [t * BIN_FIXED_CAP, +min(bin_count[t], BIN_FIXED_CAP)); 0: CSR (bin_off)
    const int *bin_count;
    const int *bin_off;
    const int *bin_list;
    const BinEntry *bin_entries; // bin_list materialised after the per-bin sort
    const int *global_list;
    const double *U;
    const double *scal;     // device: {1/dt, rhof} of this step (uploaded per step so that a captured graph stays valid)
    double *As, *Fs, *Ts, *Ct;
    double *force_torque;   // [6*n_solids], zeroed
    unsigned *pair_counts;  // [3*n_solids], zeroed
    int *slots;             // [K][n_cells] slot records (see file header)
    unsigned char *n_item;  // [n_cells] slots in use
    int2 *heavy;            // queue of (cell, solid) needing exact evaluation
    double2 *heavy_res;     // [queue] per item: (solid volume inside the cell, bits: CELL_TYPE | solid << 2)
    unsigned long long *heavy_count;
    unsigned long long *heavy_gen;   // mixed meshes: count of the items queued from the back (non-hexahedral cells)
    long long heavy_cap;
    int K;
    const unsigned char *excluded; // replay pass: [n_cells*K] 1 = pair is outside the seed's component
    StepStatus *status;
    int c_begin, c_end;     // k_final: cell range of this launch
    int cls_begin, cls_end; // k_classify: position range of this launch (the host-buffer path works slab by slab)
    const unsigned long long *heavy_start;   // k_heavy: [0] first front-queue index, [1] first back-queue item of this launch (nullptr: 0)
};

// candidate records of tile t: [bi, be) in bin_entries
__device__ __forceinline__ void bin_range(const InteractParams &P, unsigned t, int &bi, int &be) {
    if (P.bin_fixed) {
        bi = (int)t * BIN_FIXED_CAP;
        be = bi + min(__ldg(P.bin_count + t), BIN_FIXED_CAP);
    } else {
        bi = __ldg(P.bin_off + t);
        be = __ldg(P.bin_off + t + 1);
    }
}

__device__ __forceinline__ float2 cell_radius(const DevMesh &m, int c) { return m.rad_uniform ? m.rad_const : __ldg(m.cell_rad + c); }

// One member pair (solidcloud.cpp:384-390,411-421): the increments of the cell's fields and of the solid's sums.
__device__ __forceinline__ void pair_terms(const DevSolid &S, D3 cc, D3 uf, double vol, double alpha, double dtINV,
                                           D3 &fs_inc, double contrib[6]) {
    const D3 t = {S.pos[0], S.pos[1], S.pos[2]};
    const D3 om = {S.omega[0], S.omega[1], S.omega[2]};
    const D3 us = D3{S.vel[0], S.vel[1], S.vel[2]} + cross3(om, cc - t);
    const D3 f_ = alpha * (uf - us);
    const D3 t_ = cross3(cc - t, f_);
    const D3 fo = f_ * vol * dtINV;
    const D3 to = t_ * vol * dtINV;
    contrib[0] = fo.x; contrib[1] = fo.y; contrib[2] = fo.z;
    contrib[3] = to.x; contrib[4] = to.y; contrib[5] = to.z;
    fs_inc = f_ * dtINV;
}

// c = the caller's cell label (orig[position])
#ifndef FINAL_STREAM_STORES
#define FINAL_STREAM_STORES 0
#endif
__device__ __forceinline__ void store_cell(const InteractParams &P, int c, double as, D3 fs, double ts, double ct) {
#if FINAL_STREAM_STORES
    // the four fields are written once per step and not read again by this pipeline: streaming (evict-first) stores
    __stcs(P.As + c, (as < 1.0) ? as : 1.0);
    __stcs(P.Fs + 3 * (long long)c, fs.x);
    __stcs(P.Fs + 3 * (long long)c + 1, fs.y);
    __stcs(P.Fs + 3 * (long long)c + 2, fs.z);
    __stcs(P.Ts + c, ts);
    __stcs(P.Ct + c, ct);
#else
    P.As[c] = (as < 1.0) ? as : 1.0;                                               // checkAlpha, :564-570 (std::min(As,1))
    P.Fs[3 * (long long)c] = fs.x;
    P.Fs[3 * (long long)c + 1] = fs.y;
    P.Fs[3 * (long long)c + 2] = fs.z;
    P.Ts[c] = ts;
    P.Ct[c] = ct;
#endif
}

// warp-level aggregation of one member pair per lane: lanes with the same solid are reduced together and
// the 6 force/torque sums + 2 of the 3 type counters leave the warp as 8 parallel reductions.  The 8 values
// are folded (16 -> 8 -> 4 lanes keep half of the values each) so the butterfly costs 7 shuffles, not 30+.
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
        double w[8];
#pragma unroll
        for (int k = 0; k < 6; ++k) w[k] = mine ? v[k] : 0.0;
        w[6] = (mine && type == 1) ? 1.0 : 0.0;
        w[7] = (mine && type == 2) ? 1.0 : 0.0;
        const unsigned c3 = __popc(__ballot_sync(FULL, mine && type == 3));
        double a[4], b[2], c;
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = (h16 ? w[i + 4] : w[i]) + __shfl_xor_sync(FULL, h16 ? w[i] : w[i + 4], 16);
#pragma unroll
        for (int i = 0; i < 2; ++i) b[i] = (h8 ? a[i + 2] : a[i]) + __shfl_xor_sync(FULL, h8 ? a[i] : a[i + 2], 8);
        c = (h4 ? b[1] : b[0]) + __shfl_xor_sync(FULL, h4 ? b[0] : b[1], 4);
        c += __shfl_xor_sync(FULL, c, 2);
        c += __shfl_xor_sync(FULL, c, 1);
        if ((lane & 3) == 0) {
            const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if (idx < 6) atomicAdd(force_torque + 6 * (long long)s0 + idx, c);
            else if (c != 0.0) atomicAdd(pair_counts + 3 * (long long)s0 + (idx - 6), (unsigned)c);
        }
        if (lane == leader && c3) atomicAdd(pair_counts + 3 * (long long)s0 + 2, c3);
        pending &= ~grp;
    }
}

// ------------------------------------------------------------------------------------------------
// k_classify: thread per mesh cell (tile order).  The candidates of a cell are the solids binned on its tile —
// one contiguous, id-sorted run of 32-byte records that the whole warp reads at the same address (broadcast) —
// tested in fp32 against the fp32 centre copy: 16 bytes of HBM traffic per cell and no fp64 instruction
// unless a plane / tilted 2-D solid (global list) is present.
// ------------------------------------------------------------------------------------------------
template <int NT, int MINB, bool HAS_GLOBAL, bool REFINE>
__global__ void __launch_bounds__(NT, MINB) k_classify(InteractParams P) {
    const DevMesh &m = P.m;
    const unsigned FULL = 0xffffffffu;
    const int c = P.cls_begin + blockIdx.x * NT + threadIdx.x;
    const bool live = c < P.cls_end;
    const long long nC = m.n_cells;
    const int lane = threadIdx.x & 31;
    int n_item = 0, n_heavy = 0;
    if (live) {
        const float4 p = __ldg(m.cc32 + c);
        const unsigned t = __ldg(m.tile_key + c);
        int bi, be;
        bin_range(P, t, bi, be);
        const float4 *E = reinterpret_cast<const float4 *>(P.bin_entries);
        // one pre-classified candidate -> slot record
        int n_over = 0;
        auto emit = [&](int s, int qc) {
            if (qc == 0) return;
            if (n_item < P.K) {
                P.slots[(long long)n_item * nC + c] = (s << 3) | (qc == 2 ? SLOT_HEAVY : SDFIBM_CELL_ALL_INSIDE);
                n_heavy += (qc == 2);
                ++n_item;
            } else ++n_over;
        };
        // fp32 three-way test of a binned candidate (two 16-byte records) against the cell's vertex-cloud box: N2 bounds the
        // nearest vertex from below (for box cells it IS the nearest corner), F2 the farthest vertex from above; r_out / r_in
        // carry the fp32 slack
        const float4 hb = m.box_uniform ? m.box_const : __ldg(m.cell_box + c);
        auto test32 = [&](float4 e0, float4 e1) {
            const bool k3 = __float_as_int(e1.z) == KIND_3D;
            const float ax = fabsf(p.x - e0.x), ay = fabsf(p.y - e0.y), az = k3 ? fabsf(p.z - e0.z) : 0.f;
            const float hz = k3 ? hb.z : 0.f;
            const float fx = ax + hb.x, fy = ay + hb.y, fz = az + hz;
            float nx = ax - hb.x, ny = ay - hb.y, nz = az - hz;
            if (hb.w == 0.f) { nx = fmaxf(nx, 0.f); ny = fmaxf(ny, 0.f); nz = fmaxf(nz, 0.f); }
            const float N2 = nx * nx + ny * ny + nz * nz, F2 = fx * fx + fy * fy + fz * fz;
            const float ro = e0.w, ri = e1.x;
            return (N2 > ro * ro) ? 0 : ((ri > 0.f && F2 < ri * ri) ? 1 : 2);
        };
        // Second look at an undecided pair for convex analytic shapes: the shape's inside function at the 8 corners of the cell's
        // box, in fp32 with a certified error bound.  Every corner inside => every vertex inside (the vertices lie in the box and
        // the shape is convex); every corner outside a BOX cell => every vertex outside (its vertices are the corners).
        auto refine32 = [&](int s, int qc, int mode) {
            if (!REFINE || qc != 2 || mode == 0) return qc;
            const DevSolid &S = P.solids[s];
            const bool k3 = S.kind == KIND_3D;
            const float dx = p.x - S.pos32[0], dy = p.y - S.pos32[1], dz = p.z - S.pos32[2];
            const float hz = k3 ? hb.z : 0.f;
            float gmax = -3.0e38f, gmin = 3.0e38f;
            float bc[3], ex[3], ey[3], ez[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                bc[i] = S.M[3 * i] * dx + S.M[3 * i + 1] * dy + S.M[3 * i + 2] * dz + S.com32[i];
                ex[i] = S.M[3 * i] * hb.x; ey[i] = S.M[3 * i + 1] * hb.y; ez[i] = S.M[3 * i + 2] * hz;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float g = (mode == 1) ? -1.f : -3.0e38f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float b = bc[i] + ((k & 1) ? ex[i] : -ex[i]) + ((k & 2) ? ey[i] : -ey[i]) + ((k & 4) ? ez[i] : -ez[i]);
                    if (mode == 1) { const float t = b * S.rp[i]; g += t * t; }
                    else g = fmaxf(g, fabsf(b) - S.rp[i]);
                }
                gmax = fmaxf(gmax, g);
                gmin = fminf(gmin, g);
            }
            const float eps = S.eps_ref;
            if (gmax < -eps) return 1;
            if (hb.w != 0.f && gmin > ((mode == 1) ? 4.f * eps : eps)) return 0;
            return 2;
        };
        if (!HAS_GLOBAL) {
            for (; bi < be; ++bi) {
                const float4 e0 = __ldg(E + 2 * (long long)bi), e1 = __ldg(E + 2 * (long long)bi + 1);
                const int s = __float_as_int(e1.y);
                emit(s, refine32(s, test32(e0, e1), __float_as_int(e1.w) & 0xff));
            }
        } else {
            // planes / tilted 2-D solids are tested by every cell: merge the tile list and the global list in ascending solid id
            int gi = 0;
            const int ge = P.status->n_global;
            const D3 cc = ld3(m.cc, c);
            const float2 rad = cell_radius(m, c);
            while (bi < be || gi < ge) {
                float4 e0 = {0.f, 0.f, 0.f, 0.f}, e1 = {0.f, 0.f, 0.f, 0.f};
                int sb = 0x7fffffff;
                if (bi < be) {
                    e0 = __ldg(E + 2 * (long long)bi);
                    e1 = __ldg(E + 2 * (long long)bi + 1);
                    sb = __float_as_int(e1.y);
                }
                const int sg = (gi < ge) ? __ldg(P.global_list + gi) : 0x7fffffff;
                if (sb <= sg) { ++bi; emit(sb, refine32(sb, test32(e0, e1), __float_as_int(e1.w) & 0xff)); }
                else { ++gi; emit(sg, quick_class(P.solids[sg], cc, rad)); }
            }
        }
        P.n_item[c] = (unsigned char)n_item;
        if (n_over) { P.status->slot_overflow = 1; atomicMax(&P.status->slot_need, n_item + n_over); }   // the host widens the records and runs again
    }
    // ---- block-aggregated append of the queued pairs to the global queue: ONE atomic on the queue counter per CTA
    //      (a per-warp atomic serialises ~3e5 same-address operations at C4 and bounds the whole kernel).  Mixed meshes keep two
    //      queues in the one buffer — hexahedral cells from the front, the others from the back — so the warp-cooperative hex
    //      kernel sees only hexahedra; both counts ride one scan, 16 bits each ----
    const bool gen_cell = m.mixed && n_heavy > 0 && __ldg(m.hex_topo + 3 * (long long)c) == HEX_NONE;
    const int mine = gen_cell ? (n_heavy << 16) : n_heavy;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    __shared__ int s_wtot[NT / 32];
    __shared__ unsigned long long s_base, s_base_gen;
    const int warp = threadIdx.x >> 5;
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) { const int t = s_wtot[w]; s_wtot[w] = tot; tot += t; }
        const int tot_hex = tot & 0xffff, tot_gen = tot >> 16;
        s_base = tot_hex ? atomicAdd(P.heavy_count, (unsigned long long)tot_hex) : 0ull;
        s_base_gen = tot_gen ? atomicAdd(P.heavy_gen, (unsigned long long)tot_gen) : 0ull;
    }
    __syncthreads();
    if (n_heavy == 0) return;
    const int excl = s_wtot[warp] + incl - mine;
    // front queue: ascending indices; back queue: item i lives at heavy_cap - 1 - i
    long long pos = gen_cell ? P.heavy_cap - 1 - ((long long)s_base_gen + (excl >> 16)) : (long long)s_base + (excl & 0xffff);
    const long long step = gen_cell ? -1 : 1;
    for (int j = 0; j < n_item; ++j) {
        const int e = P.slots[(long long)j * nC + c];
        if (e & SLOT_HEAVY) {
            if (pos >= 0 && pos < P.heavy_cap) {
                P.heavy[pos] = make_int2(c, e >> 3);
                P.slots[(long long)j * nC + c] = ((int)pos << 3) | SLOT_HEAVY;     // the slot now points at its queue item
            }
            pos += step;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_classify4: the plain classification (no global list, no corner refinement — the C4 variant), FOUR consecutive positions per
// thread.  The candidate records of a tile are fetched once per four cells, the four tests are independent instruction streams, slot 0 of the four cells is held in registers and leaves as
// one 16-byte store (every sector of slots[0][.] written in full: no partial-sector fills), n_item as one 4-byte store, and the
// scan / queue append run once per four cells.  Queue order is unchanged (position order).  A thread whose four positions do
// not share one tile (tile runs need not be multiples of four on general meshes) walks them one by one.  Measured at C4:
// 0.225 -> 0.18 ms; with the refinement / global-list code in the loop the register count halves the occupancy and the
// one-position kernel stays faster, so those variants keep it.
// ------------------------------------------------------------------------------------------------
// k_classify's fp32 corner refinement (the same operation sequence) as a function of its own: k_classify4 calls it out of line, once
// per undecided (cell, refinable solid) pair, so that its registers are not paid four times over in the candidate loop.
__device__ __noinline__ int refine32_pair(const DevSolid &S, float4 p, float4 hb, int mode) {
    const bool k3 = S.kind == KIND_3D;
    const float dx = p.x - S.pos32[0], dy = p.y - S.pos32[1], dz = p.z - S.pos32[2];
    const float hz = k3 ? hb.z : 0.f;
    float gmax = -3.0e38f, gmin = 3.0e38f;
    float bc[3], ex[3], ey[3], ez[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        bc[i] = S.M[3 * i] * dx + S.M[3 * i + 1] * dy + S.M[3 * i + 2] * dz + S.com32[i];
        ex[i] = S.M[3 * i] * hb.x; ey[i] = S.M[3 * i + 1] * hb.y; ez[i] = S.M[3 * i + 2] * hz;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float g = (mode == 1) ? -1.f : -3.0e38f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float b = bc[i] + ((k & 1) ? ex[i] : -ex[i]) + ((k & 2) ? ey[i] : -ey[i]) + ((k & 4) ? ez[i] : -ez[i]);
            if (mode == 1) { const float t = b * S.rp[i]; g += t * t; }
            else g = fmaxf(g, fabsf(b) - S.rp[i]);
        }
        gmax = fmaxf(gmax, g);
        gmin = fminf(gmin, g);
    }
    const float eps = S.eps_ref;
    if (gmax < -eps) return 1;
    if (hb.w != 0.f && gmin > ((mode == 1) ? 4.f * eps : eps)) return 0;
    return 2;
}

// SPEC: every shape of the table is 3-D and every cell is the same axis-aligned box (C4, C5): the 2-D / 3-D selects and the clamp of
// the near-corner offsets drop out of the test.  The two sums of squares use explicit fmaf (one rounding less per term; the 4e-6
// relative slack of the radii covers either rounding sequence — tests/test_classify_bounds_cpu.py runs both).
template <int NT, int MINB, bool SPEC, bool REFINE, bool PART>
__global__ void __launch_bounds__(NT, MINB) k_classify4(InteractParams P) {
    const DevMesh &m = P.m;
    const unsigned FULL = 0xffffffffu;
    const long long nC = m.n_cells;
    const int lane = threadIdx.x & 31;
    const int c0 = P.cls_begin + (blockIdx.x * NT + threadIdx.x) * 4;
    const bool vec = ((P.cls_begin | (int)(nC & 3)) & 3) == 0;   // 16-byte accesses are aligned (uniform over the grid)
    const bool full = c0 + 3 < P.cls_end;
    int n_item[4] = {0, 0, 0, 0}, n_heavy[4] = {0, 0, 0, 0}, slot0[4] = {0, 0, 0, 0};
    int n_over[4] = {0, 0, 0, 0};
    unsigned ballmask[4] = {0u, 0u, 0u, 0u};   // bit j: slot j of the cell is a queued pair of an un-rotated sphere (front queue)
    const float4 *E = reinterpret_cast<const float4 *>(P.bin_entries);
    if (c0 < P.cls_end) {
        float4 p[4], hb[4];
        unsigned t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = min(c0 + q, P.cls_end - 1);
            p[q] = __ldg(m.cc32 + c);
            hb[q] = (SPEC || m.box_uniform) ? m.box_const : __ldg(m.cell_box + c);
        }
        if (vec && full) {
            const uint4 tk = __ldg(reinterpret_cast<const uint4 *>(m.tile_key + c0));
            t[0] = tk.x; t[1] = tk.y; t[2] = tk.z; t[3] = tk.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) t[q] = __ldg(m.tile_key + min(c0 + q, P.cls_end - 1));
        }
        // one pre-classified candidate of cell q -> slot record (slot 0 stays in a register)
        auto emit = [&](int q, int s, int qc, int ball) {
            if (qc == 0) return;
            if (n_item[q] < P.K) {
                const int rec = (s << 3) | (qc == 2 ? SLOT_HEAVY : SDFIBM_CELL_ALL_INSIDE);
                if (n_item[q] == 0) slot0[q] = rec;
                else P.slots[(long long)n_item[q] * nC + (c0 + q)] = rec;
                n_heavy[q] += (qc == 2);
                if (PART && qc == 2 && ball && n_item[q] < 32) ballmask[q] |= 1u << n_item[q];
                ++n_item[q];
            } else ++n_over[q];
        };
        auto test32 = [&](int q, float4 e0, float4 e1) {     // see k_classify
            const bool k3 = SPEC || __float_as_int(e1.z) == KIND_3D;
            const float ax = fabsf(p[q].x - e0.x), ay = fabsf(p[q].y - e0.y), az = k3 ? fabsf(p[q].z - e0.z) : 0.f;
            const float hz = k3 ? hb[q].z : 0.f;
            const float fx = ax + hb[q].x, fy = ay + hb[q].y, fz = az + hz;
            float nx = ax - hb[q].x, ny = ay - hb[q].y, nz = az - hz;
            if (!SPEC && hb[q].w == 0.f) { nx = fmaxf(nx, 0.f); ny = fmaxf(ny, 0.f); nz = fmaxf(nz, 0.f); }
            const float N2 = fmaf(nz, nz, fmaf(ny, ny, nx * nx)), F2 = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
            const float ro = e0.w, ri = e1.x;
            return (N2 > ro * ro) ? 0 : ((ri > 0.f && F2 < ri * ri) ? 1 : 2);
        };
        // SPEC: the four positions of a thread are normally four consecutive cells of one mesh line (tiles keep the caller's order
        // inside, x fastest): their y and z offsets to a candidate — and the partial sums of squares they feed — are taken once
        const bool line = SPEC && p[1].y == p[0].y && p[2].y == p[0].y && p[3].y == p[0].y && p[1].z == p[0].z && p[2].z == p[0].z && p[3].z == p[0].z;
        if (line && t[1] == t[0] && t[2] == t[0] && t[3] == t[0]) {
            int bi, be;
            bin_range(P, t[0], bi, be);
            const float4 h = m.box_const;
            for (; bi < be; ++bi) {
                const float4 e0 = __ldg(E + 2 * (long long)bi), e1 = __ldg(E + 2 * (long long)bi + 1);
                const int s = __float_as_int(e1.y);
                const float ay = fabsf(p[0].y - e0.y), az = fabsf(p[0].z - e0.z);
                const float fy = ay + h.y, fz = az + h.z, ny = ay - h.y, nz = az - h.z;
                const float Nyz = fmaf(nz, nz, ny * ny), Fyz = fmaf(fz, fz, fy * fy);
                const float ro2 = e0.w * e0.w, ri2 = e1.x * e1.x;
                const bool has_in = e1.x > 0.f;
                int qc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float ax = fabsf(p[q].x - e0.x);
                    const float fx = ax + h.x, nx = ax - h.x;
                    const float N2 = fmaf(nx, nx, Nyz), F2 = fmaf(fx, fx, Fyz);
                    qc[q] = (N2 > ro2) ? 0 : ((has_in && F2 < ri2) ? 1 : 2);
                }
                const int flags = __float_as_int(e1.w), mode = flags & 0xff, ball = (flags >> 8) & 1;
                if (REFINE && mode != 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (qc[q] == 2) qc[q] = refine32_pair(P.solids[s], p[q], hb[q], mode);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) if (full || c0 + q < P.cls_end) emit(q, s, qc[q], ball);
            }
        } else if (t[1] == t[0] && t[2] == t[0] && t[3] == t[0]) {
            int bi, be;
            bin_range(P, t[0], bi, be);
#if CLS4_PREFETCH
            float4 e0 = {0.f, 0.f, 0.f, 0.f}, e1 = e0;
            if (bi < be) { e0 = __ldg(E + 2 * (long long)bi); e1 = __ldg(E + 2 * (long long)bi + 1); }
#endif
            for (; bi < be; ++bi) {
#if CLS4_PREFETCH
                float4 f0 = e0, f1 = e1;
                if (bi + 1 < be) { f0 = __ldg(E + 2 * (long long)bi + 2); f1 = __ldg(E + 2 * (long long)bi + 3); }   // next record in flight
#else
                const float4 e0 = __ldg(E + 2 * (long long)bi), e1 = __ldg(E + 2 * (long long)bi + 1);
#endif
                const int s = __float_as_int(e1.y);
                int qc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) qc[q] = test32(q, e0, e1);
                const int flags = __float_as_int(e1.w), mode = flags & 0xff, ball = (flags >> 8) & 1;
                if (REFINE && mode != 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (qc[q] == 2) qc[q] = refine32_pair(P.solids[s], p[q], hb[q], mode);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) if (full || c0 + q < P.cls_end) emit(q, s, qc[q], ball);
#if CLS4_PREFETCH
                e0 = f0; e1 = f1;
#endif
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (c0 + q >= P.cls_end) continue;
                int bi, be;
                bin_range(P, t[q], bi, be);
                for (; bi < be; ++bi) {
                    const float4 e0 = __ldg(E + 2 * (long long)bi), e1 = __ldg(E + 2 * (long long)bi + 1);
                    int qc1 = test32(q, e0, e1);
                    const int flags = __float_as_int(e1.w), mode = flags & 0xff;
                    if (REFINE && qc1 == 2 && mode != 0) qc1 = refine32_pair(P.solids[__float_as_int(e1.y)], p[q], hb[q], mode);
                    emit(q, __float_as_int(e1.y), qc1, (flags >> 8) & 1);
                }
            }
        }
        if (vec && full) *reinterpret_cast<uchar4 *>(P.n_item + c0) = make_uchar4((unsigned char)n_item[0], (unsigned char)n_item[1], (unsigned char)n_item[2], (unsigned char)n_item[3]);
        else {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (c0 + q < P.cls_end) P.n_item[c0 + q] = (unsigned char)n_item[q];
        }
        if (n_over[0] | n_over[1] | n_over[2] | n_over[3]) {   // the host widens the records and runs again
            P.status->slot_overflow = 1;
            atomicMax(&P.status->slot_need, max(max(n_item[0] + n_over[0], n_item[1] + n_over[1]), max(n_item[2] + n_over[2], n_item[3] + n_over[3])));
        }
    }
    // ---- block-aggregated append to the queue (see k_classify): hex count in the low, general-cell count in the high 16 bits
    // (part_shapes, exact-box meshes: the low count = pairs of un-rotated spheres, front queue; the high count = all other pairs, back queue)
    int mine = 0;
    bool gen[4] = {false, false, false, false};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (PART) {
            const int nb = __popc(ballmask[q]);
            mine += nb | ((n_heavy[q] - nb) << 16);
        } else {
            gen[q] = m.mixed && n_heavy[q] > 0 && __ldg(m.hex_topo + 3 * (long long)(c0 + q)) == HEX_NONE;
            mine += gen[q] ? (n_heavy[q] << 16) : n_heavy[q];
        }
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    __shared__ int s_wtot[NT / 32];
    __shared__ unsigned long long s_base, s_base_gen;
    const int warp = threadIdx.x >> 5;
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) { const int tw = s_wtot[w]; s_wtot[w] = tot; tot += tw; }
        const int tot_hex = tot & 0xffff, tot_gen = tot >> 16;
        s_base = tot_hex ? atomicAdd(P.heavy_count, (unsigned long long)tot_hex) : 0ull;
        s_base_gen = tot_gen ? atomicAdd(P.heavy_gen, (unsigned long long)tot_gen) : 0ull;
    }
    __syncthreads();
    if (c0 >= P.cls_end) return;
    if (mine != 0) {
        const int excl = s_wtot[warp] + incl - mine;
        long long pos_hex = (long long)s_base + (excl & 0xffff);
        long long pos_gen = P.heavy_cap - 1 - ((long long)s_base_gen + (excl >> 16));   // back queue: item i lives at heavy_cap - 1 - i
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (n_heavy[q] == 0) continue;
            const int c = c0 + q;
            for (int j = 0; j < n_item[q]; ++j) {
                const int e = (j == 0) ? slot0[q] : P.slots[(long long)j * nC + c];
                if (e & SLOT_HEAVY) {
                    const bool back = PART ? !(j < 32 && ((ballmask[q] >> j) & 1u)) : gen[q];
                    const long long pos = back ? pos_gen : pos_hex;
                    if (pos >= 0 && pos < P.heavy_cap) {
                        P.heavy[pos] = make_int2(c, e >> 3);
                        const int rec = ((int)pos << 3) | SLOT_HEAVY;     // the slot now points at its queue item
                        if (j == 0) slot0[q] = rec;
                        else P.slots[(long long)j * nC + c] = rec;
                    }
                    if (back) --pos_gen; else ++pos_hex;
                }
            }
        }
    }
    if (vec && full) *reinterpret_cast<int4 *>(P.slots + c0) = make_int4(slot0[0], slot0[1], slot0[2], slot0[3]);
    else {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (c0 + q < P.cls_end && n_item[q] > 0) P.slots[c0 + q] = slot0[q];
    }
}

// ------------------------------------------------------------------------------------------------
// k_heavy: exact evaluation of one (cell, solid) item
// ------------------------------------------------------------------------------------------------
// General polyhedra: cell-local arrays in local memory, CSR connectivity.
template <bool PROG>
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
    const bool ident = quat_is_identity(q);
    for (int k = 0; k < nv; ++k) {
        vid[k] = __ldg(m.cp + pb + k);
        pts[k] = ld3(m.points, vid[k]);
        double ph;
        n_in += shape_eval<true, PROG>(sh.s, world2local_sel(q, t, pts[k], ident), ph, P.ops) ? 1 : 0;
        phi[k] = ph;
    }
    type_out = 0;
    vol_out = 0.0;
    if (n_in == 0) return;
    if (n_in == nv) { type_out = SDFIBM_CELL_ALL_INSIDE; return; }
    double dummy;
    type_out = shape_eval<false, PROG>(sh.s, world2local_sel(q, t, ld3(m.cc, c), ident), dummy, P.ops) ? SDFIBM_CELL_CENTER_INSIDE : SDFIBM_CELL_CENTER_OUTSIDE;
    vol_out = cell_solid_volume(m, c, vid, pts, phi, nv);
}

// Hexahedral fast path, warp-cooperative.  Lane = one (cell, solid) item; consecutive lanes are mostly
// consecutive cells of one solid along a mesh line, whose "low" vertex quad {0,2,4,6} (cellPoints order)
// is the previous lane's "high" quad {1,3,5,7}.  Every lane evaluates its own high quad (4 rounds); the low
// quads are evaluated only for run starts, spread over all lanes in extra rounds.  A vertex shared by two
// lanes is therefore transformed and evaluated once — the SAME operation sequence either lane would run, so
// results are bit-identical to the per-cell evaluation.  Coordinates, phi and inside flags are staged in
// transposed shared memory (column = lane's thread) so that the data-dependent face->slot indexing of the
// volume phase is bank-conflict free.
struct HeavySmem {
    double px[8 * TPB], py[8 * TPB], pz[8 * TPB], phi[8 * TPB];
    double cx[TPB], cy[TPB], cz[TPB];   // cell centre of the lane's item
    unsigned char in[8 * TPB];
};

// 8-byte asynchronous global -> shared copy: the coordinates of a lane's 8 vertices (and its cell centre) are all in flight
// at once without holding registers, and land directly in the transposed layout the evaluation reads.
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void stage_point(HeavySmem &sm, const double *__restrict__ points, int v, int slot, int col) {
    const double *g = points + 3 * (long long)v;
    cp_async8(&sm.px[slot * TPB + col], g);
    cp_async8(&sm.py[slot * TPB + col], g + 1);
    cp_async8(&sm.pz[slot * TPB + col], g + 2);
}

// evaluate the staged vertex (slot_a, col_a); dup: the same vertex is slot_b of the neighbouring lane's cell
template <bool PROG>
__device__ __forceinline__ void eval_staged_vertex(HeavySmem &sm, const DevShape &sh, const sdfibm_sdf_op_t *ops, DQ q, D3 t, bool ident,
                                                   int slot_a, int col_a, bool dup, int slot_b, int col_b) {
    const D3 p = {sm.px[slot_a * TPB + col_a], sm.py[slot_a * TPB + col_a], sm.pz[slot_a * TPB + col_a]};
    double ph;
    const bool in = shape_eval<true, PROG>(sh.s, world2local_sel(q, t, p, ident), ph, ops);
    sm.phi[slot_a * TPB + col_a] = ph;
    sm.in[slot_a * TPB + col_a] = in ? 1 : 0;
    if (dup) {
        sm.phi[slot_b * TPB + col_b] = ph;
        sm.in[slot_b * TPB + col_b] = in ? 1 : 0;
    }
}

#define HEAVY_CTAS_PER_SM 5
template <int CTAS, bool PROG>
__global__ void __launch_bounds__(TPB, CTAS) k_heavy_hex(InteractParams P) {
    __shared__ HeavySmem sm;
    const DevMesh &m = P.m;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    const long long n = min((long long)*P.heavy_count, P.heavy_cap - (m.mixed ? (long long)*P.heavy_gen : 0));
    const long long q0 = P.heavy_start ? (long long)*P.heavy_start : 0;
    for (long long k0 = q0 + (long long)blockIdx.x * TPB + wbase; k0 < n; k0 += (long long)gridDim.x * TPB) {
        const long long k = k0 + lane;
        const bool valid = k < n;
        int c = -1, s = -1;
        int vid[8] = {-1, -2, -3, -4, -5, -6, -7, -8};
        DQ q = {1.0, {0.0, 0.0, 0.0}};
        D3 t = {0.0, 0.0, 0.0};
        int shape_idx = 0;
        unsigned tw0 = 0, tw1 = 0, tw2 = 0;
        int2 f01 = {0, 0}, f23 = {0, 0}, f45 = {0, 0};
        if (valid) {
            const int2 it = __ldg(P.heavy + k);
            c = it.x;
            s = it.y;
            if (m.is_hex) {
                const int4 *cp4 = reinterpret_cast<const int4 *>(m.cp + 8 * (long long)c);
                const int4 va = __ldg(cp4), vb = __ldg(cp4 + 1);
                vid[0] = va.x; vid[1] = va.y; vid[2] = va.z; vid[3] = va.w;
                vid[4] = vb.x; vid[5] = vb.y; vid[6] = vb.z; vid[7] = vb.w;
            } else {   // a hexahedron of a mixed mesh: CSR offsets, no alignment guarantee
                const int *cpp = m.cp + __ldg(m.cp_off + c);
#pragma unroll
                for (int j = 0; j < 8; ++j) vid[j] = __ldg(cpp + j);
            }
            const DevSolid &S = P.solids[s];
            q = {S.q[0], {S.q[1], S.q[2], S.q[3]}};
            t = {S.pos[0], S.pos[1], S.pos[2]};
            shape_idx = S.shape;
            // issued early: consumed by the volume phase
            tw0 = __ldg(m.hex_topo + 3 * (long long)c);
            tw1 = __ldg(m.hex_topo + 3 * (long long)c + 1);
            tw2 = __ldg(m.hex_topo + 3 * (long long)c + 2);
            if (m.is_hex) {
                const int2 *cf2 = reinterpret_cast<const int2 *>(m.cf + 6 * (long long)c);
                f01 = __ldg(cf2); f23 = __ldg(cf2 + 1); f45 = __ldg(cf2 + 2);
            } else {
                const int *cfp = m.cf + __ldg(m.cf_off + c);
                f01 = make_int2(__ldg(cfp), __ldg(cfp + 1)); f23 = make_int2(__ldg(cfp + 2), __ldg(cfp + 3)); f45 = make_int2(__ldg(cfp + 4), __ldg(cfp + 5));
            }
        }
        // does my low quad coincide with the previous lane's high quad (same solid)?
        const int ls = __shfl_up_sync(FULL, s, 1);
        const int l1 = __shfl_up_sync(FULL, vid[1], 1), l3 = __shfl_up_sync(FULL, vid[3], 1);
        const int l5 = __shfl_up_sync(FULL, vid[5], 1), l7 = __shfl_up_sync(FULL, vid[7], 1);
        const bool shared_left = valid && lane > 0 && ls == s && l1 == vid[0] && l3 == vid[2] && l5 == vid[4] && l7 == vid[6];
        const unsigned sharedmask = __ballot_sync(FULL, shared_left);
        const unsigned startmask = __ballot_sync(FULL, valid) & ~sharedmask;
        const bool right_shares = lane < 31 && ((sharedmask >> (lane + 1)) & 1u);
        const DevShape &sh = P.shapes[shape_idx];
        // warp-uniform: every solid of this warp's items still has the identity orientation (world2local == p - t)
        const bool ident = __all_sync(FULL, !valid || quat_is_identity(q));
        // stage the coordinates of my 8 vertices and my cell centre: 27 asynchronous 8-byte copies, one memory latency
        if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) stage_point(sm, m.points, vid[j], j, tid);
            const double *g = m.cc + 3 * (long long)c;
            cp_async8(&sm.cx[tid], g);
            cp_async8(&sm.cy[tid], g + 1);
            cp_async8(&sm.cz[tid], g + 2);
        }
        cp_async_wait_all();
        __syncwarp();
        // rounds 0..3: my own high quad (also the next lane's low quad when it shares)
        if (valid) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
                eval_staged_vertex<PROG>(sm, sh, P.ops, q, t, ident, 2 * v + 1, tid, right_shares, 2 * v, tid + 1);
        }
        // extra rounds: low quads of the run starts, 4 vertices each, spread over all lanes
        const int nwork = 4 * __popc(startmask);
        for (int base = 0; base < nwork; base += 32) {
            const int u = base + lane;
            const bool work = u < nwork;
            const int src = work ? (int)__fns(startmask, 0, (u >> 2) + 1) : 0;
            const int s_src = __shfl_sync(FULL, s, src);
            if (work) {
                const int kk = 2 * (u & 3);
                const int col = wbase + src;
                const DevSolid &S2 = P.solids[s_src];
                const DQ q2 = {S2.q[0], {S2.q[1], S2.q[2], S2.q[3]}};
                const D3 t2 = {S2.pos[0], S2.pos[1], S2.pos[2]};
                eval_staged_vertex<PROG>(sm, P.shapes[S2.shape], P.ops, q2, t2, ident, kk, col, false, 0, 0);
            }
        }
        __syncwarp();
        if (valid) {
            int type = 0;
            double volume = 0.0;
            int n_in = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) n_in += sm.in[j * TPB + tid];
            if (n_in == 8) type = SDFIBM_CELL_ALL_INSIDE;
            else if (n_in != 0) {
                double dummy;
                type = shape_eval<false, PROG>(sh.s, world2local_sel(q, t, D3{sm.cx[tid], sm.cy[tid], sm.cz[tid]}, ident), dummy, P.ops) ? SDFIBM_CELL_CENTER_INSIDE : SDFIBM_CELL_CENTER_OUTSIDE;
                auto PT = [&](int l) { return D3{sm.px[l * TPB + tid], sm.py[l * TPB + tid], sm.pz[l * TPB + tid]}; };
                auto PH = [&](int l) { return sm.phi[l * TPB + tid]; };
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
#pragma unroll 1
                for (int f = 0; f < 6; ++f) {
                    const unsigned w = (f < 2) ? tw0 : (f < 4) ? tw1 : tw2;
                    const unsigned nib = (w >> (16 * (f & 1))) & 0xffffu;
                    const int l[4] = {(int)(nib & 0xf), (int)((nib >> 4) & 0xf), (int)((nib >> 8) & 0xf), (int)((nib >> 12) & 0xf)};
                    const double ph[4] = {PH(l[0]), PH(l[1]), PH(l[2]), PH(l[3])};
                    const int npos = (ph[0] > 0) + (ph[1] > 0) + (ph[2] > 0) + (ph[3] > 0);
                    if (npos == 4) continue;                                        // eps_f = 0: adds +0.0 (:107-108)
                    const int face = (f == 0) ? f01.x : (f == 1) ? f01.y : (f == 2) ? f23.x : (f == 3) ? f23.y : (f == 4) ? f45.x : f45.y;
                    // face record: Cf.xyz, Sf.xyz, |Sf|, pad — four 16-byte loads issued before the area math
                    const double2 *fr = m.face_rec + 4 * (long long)face;
                    const double2 r0 = __ldg(fr), r1 = __ldg(fr + 1), r2 = __ldg(fr + 2), r3 = __ldg(fr + 3);
                    double eps_f = 1.0;                                             // all phi <= 0 (:109-110)
                    if (npos != 0) {
                        const D3 A = PT(l[0]);                                      // calcFaceArea (:74-96)
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
                            const D3 O = PT(l[e]), A2 = PT(l[(e + 1) & 3]);
                            area += fabs(0.5 * mag3(cross3(A2 - O, fap - O))) * lf;   // a zero fraction adds +0.0
                        }
                        eps_f = area / r3.x;
                    }
                    const D3 Cf = {r0.x, r0.y, r1.x}, Sf = {r1.y, r2.x, r2.y};
                    volume += (1.0 / 3.0) * eps_f * fabs(dot3(apex - Cf, Sf));
                }
            }
            P.heavy_res[k] = make_double2(volume, __longlong_as_double((long long)(type | (s << 2))));   // type 0: no vertex inside -> not a member
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// k_heavy_box: the exact evaluation on exact-box meshes (DevMesh::box_exact).  Lane = one (cell, solid) item.  The cell is six
// coordinates; its corners are addressed by a 3-bit code (bit0 x-hi, bit1 y-hi, bit2 z-hi).  The vertex predicates / signed
// distances, the two kinds of apex and every difference that can cancel (apex - face plane, face apex - edge) run the reference's
// own operation sequence un-contracted — the apex weights |phiA| / (SMALL + |phiA| + |phiB|) keep the IEEE division — so the lists
// are bit-exact and the fractions keep their RELATIVE accuracy on sliver cells.
// What the box buys (geometrictools.cpp:74-116 on an axis-aligned rectangle): the triangle (O, A, face apex) of calcFaceArea
// has the area 0.5 * |edge| * |offset of the apex perpendicular to the edge| — its cross product has ONE non-zero component,
// whose square root is exact — and eps_f * |(apex - Cf) . Sf| = area * |apex_a - Cf_a| because Sf is exactly (0, 0, +-|Sf|):
// no cross products, no square roots, no face records.
// Every lane runs ONE instruction stream.  The 12 edge fractions of the box are taken once (each edge belongs to two faces;
// calcLineFraction is symmetric in its arguments, and out - in == |a| + |b| bit for bit when the signs differ; they only scale
// positive terms, so a 1-ulp reciprocal sequence serves them), the six face terms are computed unconditionally and selected at
// the end (a boundary cell has 3-5 cut faces, and a warp executes the union of its lanes' branches anyway).
// ------------------------------------------------------------------------------------------------
#ifndef BOX_CTAS_PER_SM
#define BOX_CTAS_PER_SM 6
#endif
#ifndef BOX_PREFETCH
#define BOX_PREFETCH 2   // C4 k_heavy_box: 0.387 ms without, 0.390 with the register prefetch, 0.378 with the cp.async one
#endif
#ifndef BOX_EDGE_COMPACT
#define BOX_EDGE_COMPACT 1
#endif
#ifndef BOX_VUNROLL
#define BOX_VUNROLL 2
#endif
#ifndef BOX_EUNROLL
#define BOX_EUNROLL 4
#endif
#ifndef BOX_FUNROLL
#define BOX_FUNROLL 2   // C4 k_heavy_box after the edge compaction: 0.378 (1) / 0.352 (2) / 0.372 (3) / 0.347 ms (6); C5 at 256^3 (mixed shapes): 0.597 (2) / 0.638 (6)
#endif
#define PRAGMA_UNROLL_(n) _Pragma(#n)
#define PRAGMA_UNROLL(n) PRAGMA_UNROLL_(unroll n)

// calcLineFraction (:13-23), branch-free
__device__ __forceinline__ double edge_fraction(double pa, double pb) {
    const bool ta = pa > 0, tb = pb > 0;
    const double den = fabs(pa) + fabs(pb);      // == out - in
    const double num = ta ? fabs(pb) : fabs(pa);  // == -in
    double q;
    if (den < 1e-290) q = num / den;   // (planes only: both ends within 1e-290 of the surface; the seed below would flush)
    else {
        // num / den to within 1 ulp: MUFU seed (2^-20), two Newton steps, one residual correction
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
        r = fma(r, fma(-den, r, 1.0), r);
        r = fma(r, fma(-den, r, 1.0), r);
        q = num * r;
        q = fma(fma(-den, q, num), r, q);
    }
    return (ta && tb) ? 0.0 : ((ta || tb) ? q : 1.0);
}

// Per-lane scratch in shared memory, column = the lane's thread (bank = thread: conflict-free whatever the row).  The mesh's
// own vertex / face-loop order (btopo) makes every corner, edge and axis index of the volume phase a run-time value; rows
// addressed by a computed index cost one LDS where a register file would cost a chain of selects — and the face / edge / vertex
// loops stay rolled, so the whole kernel is ~1k instructions and lives in the instruction cache.
struct BoxSmem {
    double phi[8 * TPB];    // row = corner code
    double lohi[6 * TPB];   // row = 2 * axis + side
    double ef[12 * TPB];    // row = 4 * axis + (bit of axis+1) + 2 * (bit of axis+2): edge along `axis` from the corner with those bits
    double apex[3 * TPB];
};

template <bool PLANE_FROM_MESH, bool PROG, bool PART>
__global__ void __launch_bounds__(TPB, BOX_CTAS_PER_SM) k_heavy_box(InteractParams P) {
    __shared__ BoxSmem sm;
    const DevMesh &m = P.m;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    double *phi = sm.phi + tid, *lohi = sm.lohi + tid, *ef = sm.ef + tid, *apx = sm.apex + tid;
    // The items of this launch as ONE index range: first the front queue (ascending positions; with part_shapes the pairs of
    // un-rotated spheres), then the back queue (item i at heavy_cap - 1 - i; every other pair) — so all warps but the one that
    // straddles the seam are uniform in the corner path they take below.
    const long long qf = P.heavy_start ? (long long)P.heavy_start[0] : 0, qb = (PART && P.heavy_start) ? (long long)P.heavy_start[1] : 0;
    const long long n_front = max(min((long long)*P.heavy_count, P.heavy_cap) - qf, 0ll);
    const long long n_tot = n_front + (PART ? max(min((long long)*P.heavy_gen, P.heavy_cap) - qb, 0ll) : 0ll);
    auto pos_of = [&](long long i) { return (!PART || i < n_front) ? qf + i : P.heavy_cap - 1 - (qb + (i - n_front)); };
    // the queue item of the NEXT batch is requested while this one is evaluated — BOX_PREFETCH 1: 8 bytes in two registers,
    // 2: an 8-byte cp.async into the lane's shared-memory slot (no registers) — so the cell / solid records of a batch are one
    // dependent level away instead of two
#if BOX_PREFETCH == 1
    int2 it_next = make_int2(0, 0);
    {
        const long long jf = (long long)blockIdx.x * TPB + wbase + lane;
        if (jf < n_tot) it_next = __ldg(P.heavy + pos_of(jf));
    }
#elif BOX_PREFETCH == 2
    __shared__ double s_item[TPB];
    {
        const long long jf = (long long)blockIdx.x * TPB + wbase + lane;
        if (jf < n_tot) cp_async8(&s_item[tid], reinterpret_cast<const double *>(P.heavy + pos_of(jf)));
    }
#endif
    for (long long i0 = (long long)blockIdx.x * TPB + wbase; i0 < n_tot; i0 += (long long)gridDim.x * TPB) {
        const long long i = i0 + lane;
        const bool valid = i < n_tot;
        const long long k = valid ? pos_of(i) : 0;       // the item's queue position: its result goes to heavy_res[k]
        int c = 0, s = 0;
#if BOX_PREFETCH == 1
        c = it_next.x; s = it_next.y;
        {
            const long long jn = i + (long long)gridDim.x * TPB;
            if (jn < n_tot) it_next = __ldg(P.heavy + pos_of(jn));
        }
#elif BOX_PREFETCH == 2
        cp_async_wait_all();
        if (valid) { const long long it = __double_as_longlong(s_item[tid]); c = (int)(it & 0xffffffffll); s = (int)(it >> 32); }
        {
            const long long jn = i + (long long)gridDim.x * TPB;
            if (jn < n_tot) cp_async8(&s_item[tid], reinterpret_cast<const double *>(P.heavy + pos_of(jn)));
        }
#else
        if (valid) { const int2 it = __ldg(P.heavy + k); c = it.x; s = it.y; }
#endif
        DQ q = {1.0, {0.0, 0.0, 0.0}};
        D3 t = {0.0, 0.0, 0.0};
        int shape_idx = 0;
        uint2 tw = {0u, 0u};
        D3 cc = {0, 0, 0};
        if (valid) {
            const double2 *b2 = reinterpret_cast<const double2 *>(m.box6 + 6 * (long long)c);
            const double2 b0 = __ldg(b2), b1 = __ldg(b2 + 1), b3 = __ldg(b2 + 2);
            lohi[0 * TPB] = b0.x; lohi[2 * TPB] = b0.y; lohi[4 * TPB] = b1.x;      // lo.xyz
            lohi[1 * TPB] = b1.y; lohi[3 * TPB] = b3.x; lohi[5 * TPB] = b3.y;      // hi.xyz
            tw = __ldg(reinterpret_cast<const uint2 *>(m.btopo) + c);
            cc = ld3(m.cc, c);
            const DevSolid &S = P.solids[s];
            q = {S.q[0], {S.q[1], S.q[2], S.q[3]}};
            t = {S.pos[0], S.pos[1], S.pos[2]};
            shape_idx = S.shape;
        }
        const DevShape &sh = P.shapes[shape_idx];
        const bool ident = __all_sync(FULL, !valid || quat_is_identity(q));
        int n_in = 0;
        unsigned pm = 0;   // bit `code`: phi of that corner > 0
        // un-rotated balls (every solid before it turns; C4): the corner coordinates are separable, so the body-frame components
        // com + (p - t) and their squares are taken once per lattice plane (6 + 6 + 6 operations instead of 8 x 9) — the same
        // expressions, in the same association, as the generic path evaluates per vertex (device_math.cuh SPHERE: bit-identical)
        const bool ball_fast = ident && __all_sync(FULL, !valid || sh.s.tag == SDFIBM_SHAPE_SPHERE);
        if (ball_fast) {
            if (valid) {
                double sq[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const int d = r >> 1;
                    const double td = d == 0 ? t.x : (d == 1 ? t.y : t.z);
                    const double b = sh.s.com[d] + (lohi[r * TPB] - td);
                    sq[r] = b * b;
                }
                const double rad = sh.s.p[0], rad2 = sh.s.p[1];
#pragma unroll
                for (int code = 0; code < 8; ++code) {
                    const double m2 = sq[code & 1] + sq[2 + ((code >> 1) & 1)] + sq[4 + ((code >> 2) & 1)];
                    const double phv = sdf_filter(sqrt(m2) - rad);
                    n_in += (m2 < rad2) ? 1 : 0;
                    pm |= (phv > 0 ? 1u : 0u) << code;
                    phi[code * TPB] = phv;
                }
            }
        } else if (valid) {
PRAGMA_UNROLL(BOX_VUNROLL)
            for (int code = 0; code < 8; ++code) {
                const D3 p = {lohi[(code & 1) * TPB], lohi[(2 + ((code >> 1) & 1)) * TPB], lohi[(4 + ((code >> 2) & 1)) * TPB]};
                double phv;
                n_in += shape_eval<true, PROG>(sh.s, world2local_sel(q, t, p, ident), phv, P.ops) ? 1 : 0;
                pm |= (phv > 0 ? 1u : 0u) << code;
                phi[code * TPB] = phv;
            }
        }
        if (valid) {
            int type = 0;
            double volume = 0.0;
            if (n_in == 8) type = SDFIBM_CELL_ALL_INSIDE;
            else if (n_in != 0) {
                double dummy;
                type = shape_eval<false, PROG>(sh.s, world2local_sel(q, t, cc, ident), dummy, P.ops) ? SDFIBM_CELL_CENTER_INSIDE : SDFIBM_CELL_CENTER_OUTSIDE;
                // cell apex over cellPoints() order (geometrictools.cpp:25-45,56-58)
                {
                    const int cA = tw.x & 7;
                    const double phiA = phi[cA * TPB];
                    int cB = cA;
                    double phiB = 0.0;
#pragma unroll 1
                    for (int i = 1; i < 8; ++i) {
                        cB = (tw.x >> (3 * i)) & 7;
                        phiB = phi[cB * TPB];
                        if (phiA * phiB <= 0) break;
                    }
                    const double w = fabs(phiA) / (SDF_SMALL + fabs(phiA) + fabs(phiB));
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double A = lohi[(2 * d + ((cA >> d) & 1)) * TPB], B = lohi[(2 * d + ((cB >> d) & 1)) * TPB];
                        apx[d * TPB] = A - w * (A - B);
                    }
                    if (m.two_d) apx[2 * TPB] = 0.0;
                }
                // the 12 edge fractions, each taken once (every edge belongs to two faces).  Row e = 4 * axis + ... starts at the corner
                // EDGE_BASE[e] = {0,2,4,6 | 0,4,1,5 | 0,1,2,3}.
#if BOX_EDGE_COMPACT
                // An edge whose ends lie on one side has the fraction 0 (both phi > 0) or 1 by definition (:14-17): stored straight from
                // the sign mask.  Only CUT edges need the quotient, and a cell has 3-6 of them (a warp: the maximum over its lanes), so
                // each lane walks the set bits of its own 12-bit cut mask instead of all 12 rows.
                {
                    const unsigned x = pm ^ (pm >> 1), y = pm ^ (pm >> 2), z = pm ^ (pm >> 4);
                    unsigned cm = (x & 1u) | ((x >> 1) & 2u) | ((x >> 2) & 4u) | ((x >> 3) & 8u)            // axis 0: corners 0,2,4,6
                                  | ((y & 1u) << 4) | (((y >> 4) & 1u) << 5) | (((y >> 1) & 1u) << 6) | (((y >> 5) & 1u) << 7)   // axis 1: 0,4,1,5
                                  | ((z & 15u) << 8);                                                                         // axis 2: 0,1,2,3
#pragma unroll
                    for (int e = 0; e < 12; ++e) {
                        const int base = e < 4 ? 2 * e : (e < 8 ? ((e & 1) << 2) | ((e >> 1) & 1) : e - 8);
                        ef[e * TPB] = ((pm >> base) & 1u) ? 0.0 : 1.0;
                    }
                    while (cm) {
                        const int e = __ffs(cm) - 1;
                        cm &= cm - 1;
                        const int d = e >> 2;
                        const int base = e >= 8 ? e - 8 : (int)((0x51406420u >> (4 * e)) & 7u);
                        ef[e * TPB] = edge_fraction(phi[base * TPB], phi[(base | (1 << d)) * TPB]);
                    }
                }
#else
PRAGMA_UNROLL(BOX_EUNROLL)
                for (int e = 0; e < 12; ++e) {
                    const int d = e >> 2, u = (d == 2) ? 0 : d + 1, v = (d == 0) ? 2 : d - 1;   // u = (d+1)%3, v = (d+2)%3
                    const int base = ((e & 1) << u) | (((e >> 1) & 1) << v);
                    ef[e * TPB] = edge_fraction(phi[base * TPB], phi[(base | (1 << d)) * TPB]);
                }
#endif
                // six pyramid terms 1/3 eps_f |(apex - Cf) . Sf| (geometrictools.cpp:61-70,98-116), face = (axis a, side sd)
                const double *cfa = PLANE_FROM_MESH ? m.cfa6 + 6 * (long long)c : nullptr;
PRAGMA_UNROLL(BOX_FUNROLL)
                for (int f = 0; f < 6; ++f) {
                    const int a = f >> 1, sd = f & 1, u = (a == 2) ? 0 : a + 1, v = (a == 0) ? 2 : a - 1;
                    // the face loop of the mesh: first vertex at in-plane corner (iu0, iv0), second vertex along u (dir 0) or v (dir 1)
                    const unsigned tq = (tw.y >> (3 * f)) & 7u;
                    const int iu0 = tq & 1, iv0 = (tq >> 1) & 1, dir = (tq >> 2) & 1;
                    const int k0 = (sd << a) | (iu0 << u) | (iv0 << v);
                    const int su = 1 << u, sv = 1 << v;
                    const int k1 = k0 ^ (dir ? sv : su), k2 = k0 ^ su ^ sv, k3 = k0 ^ (dir ? su : sv);
                    const double phA = phi[k0 * TPB], ph1 = phi[k1 * TPB], ph2 = phi[k2 * TPB], ph3 = phi[k3 * TPB];
                    const int npos = __popc(pm & (unsigned)((0xF00FCC33AA55ull >> (8 * f)) & 0xffu));   // corners of face f with phi > 0
                    const double lu = lohi[(2 * u) * TPB], hu = lohi[(2 * u + 1) * TPB], lv = lohi[(2 * v) * TPB], hv = lohi[(2 * v + 1) * TPB];
                    const double plane = PLANE_FROM_MESH ? __ldg(cfa + f) : lohi[f * TPB];
                    const double height = fabs(apx[a * TPB] - plane);                 // |(apex - Cf) . Sf| / |Sf|
                    const double Lu = hu - lu, Lv = hv - lv;
                    // calcFaceArea (:74-96): the face apex from the loop's first vertex and the first vertex across the surface
                    const bool x1 = phA * ph1 <= 0, x2 = phA * ph2 <= 0;
                    const int kB = x1 ? k1 : (x2 ? k2 : k3);
                    const double phB = x1 ? ph1 : (x2 ? ph2 : ph3);
                    const double w = fabs(phA) / (SDF_SMALL + fabs(phA) + fabs(phB));
                    const double Au = iu0 ? hu : lu, Av = iv0 ? hv : lv;
                    const double Bu = ((kB >> u) & 1) ? hu : lu, Bv = ((kB >> v) & 1) ? hv : lv;
                    const double fu = Au - w * (Au - Bu), fv = Av - w * (Av - Bv);
                    // four triangles: edge length x perpendicular offset of the face apex, weighted by the edge's wet fraction
                    const double da = fabs(fu - lu), db = fabs(fu - hu), dc = fabs(fv - lv), dd = fabs(fv - hv);
                    const double eu0 = ef[(4 * u + 2 * sd) * TPB], eu1 = ef[(4 * u + 1 + 2 * sd) * TPB];   // along u at v = lo / hi
                    const double ev0 = ef[(4 * v + sd) * TPB], ev1 = ef[(4 * v + sd + 2) * TPB];           // along v at u = lo / hi
                    const double cut = 0.5 * fma(Lu, fma(dc, eu0, dd * eu1), Lv * fma(da, ev0, db * ev1));
                    const double area = (npos == 0) ? Lu * Lv : cut;                      // eps_f = 1 (:109-110)
                    const double term = (1.0 / 3.0) * area * height;
                    volume += (npos == 4) ? 0.0 : term;                                   // eps_f = 0 (:107-108)
                }
            }
            P.heavy_res[k] = make_double2(volume, __longlong_as_double((long long)(type | (s << 2))));
        }
    }
}

template <bool PROG>
__global__ void __launch_bounds__(TPB) k_heavy_general(InteractParams P) {
    // all-polyhedral meshes: the front queue; mixed meshes: the back queue (item i at heavy_cap - 1 - i)
    const bool back = P.m.mixed != 0;
    const long long n = back ? (long long)*P.heavy_gen : min((long long)*P.heavy_count, P.heavy_cap);
    const long long q0 = P.heavy_start ? (long long)P.heavy_start[back ? 1 : 0] : 0;
    const long long front = back ? (long long)*P.heavy_count : 0;   // the back queue may not run into the front one
    for (long long i = q0 + (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += (long long)gridDim.x * TPB) {
        const long long k = back ? P.heavy_cap - 1 - i : i;
        if (k < front) continue;
        const int2 it = __ldg(P.heavy + k); // (cell, solid)
        int type;
        double v;
        heavy_eval_general<PROG>(P, it.x, it.y, type, v);
        P.heavy_res[k] = make_double2(v, __longlong_as_double((long long)(type | (it.y << 2))));
    }
}

// ------------------------------------------------------------------------------------------------
// k_final: thread per cell, every output sector written exactly once and in full.
// ------------------------------------------------------------------------------------------------
#ifndef FINAL_NT
#define FINAL_NT 64     // 2-warp CTAs: a CTA's slot frees as soon as its slowest warp is done (256-thread CTAs: 0.440 ms, 64: 0.403 ms at C4)
#endif
// one cell of the pass, given its first-level records: caller label oc, slot count n, slot 0
__device__ __forceinline__ void final_cell(const InteractParams &P, int c, bool live, int oc, int n, int e0) {
    const DevMesh &m = P.m;
    const long long nC = m.n_cells;
    const unsigned FULL = 0xffffffffu;
    int nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));

    double as = 0.0, ts = 0.0, ct = 0.0;
    D3 fs = {0.0, 0.0, 0.0};
    if (nmax > 0) {
        const double dtINV = __ldg(P.scal);
        D3 cc = {0, 0, 0}, uf = {0, 0, 0};
        double vol = 1.0;
        if (n > 0) {
            cc = ld3(m.cc, c);
            uf = ld3(P.U, oc);
            vol = m.V_uniform ? m.V_const : __ldg(m.V + c);
        }
        for (int j = 0; j < nmax; ++j) {
            bool have = false;
            int s = -1, type = 0;
            double contrib[6] = {0, 0, 0, 0, 0, 0};
            if (j < n) {
                const int e = (j == 0) ? e0 : P.slots[(long long)j * nC + c];
                s = e >> 3;
                type = e & 3;
                double v = 0.0;
                if (e & SLOT_HEAVY) {                                           // the slot points at its queue item
                    const double2 r = P.heavy_res[e >> 3];
                    const int bits = (int)__double_as_longlong(r.y);
                    type = bits & 3;
                    s = bits >> 2;
                    v = r.x;
                    P.slots[(long long)j * nC + c] = (s << 3) | SLOT_HEAVY | type;   // final record
                }
                const bool skip = P.excluded && P.excluded[(long long)c * P.K + j]; // replay: outside the seed's component
                if (type != 0 && !skip) {
                    const double alpha = (type == SDFIBM_CELL_ALL_INSIDE) ? 1.0 : v / vol;   // solidcloud.cpp:408-410
                    D3 fi;
                    pair_terms(P.solids[s], cc, uf, vol, alpha, dtINV, fi, contrib);   // :384-390,411-421
                    as += alpha;
                    fs = fs + fi;
                    ts += alpha;
                    ct = (type == SDFIBM_CELL_ALL_INSIDE) ? (double)(s + 4) : (double)type;   // :376-382, last writer wins
                    have = true;
                }
            }
            if (__any_sync(FULL, have)) warp_accumulate(have, s, type, contrib, P.force_torque, P.pair_counts);
        }
    }
    if (live) store_cell(P, oc, as, fs, ts, ct);
}

template <int MINB>
__global__ void __launch_bounds__(FINAL_NT, MINB) k_final(InteractParams P) {
    const int c = P.c_begin + blockIdx.x * blockDim.x + threadIdx.x;   // [c_begin, c_end): one chunk of the cell range
    const bool live = c < P.c_end;
    const int oc = live ? __ldg(P.m.orig + c) : 0;   // caller's cell label: U is read and the fields are written there
    const int n = live ? (int)P.n_item[c] : 0;
    const int e0 = live ? P.slots[c] : 0;       // slot 0, fetched together with n_item (meaningless when n == 0)
    final_cell(P, c, live, oc, n, e0);
}

// ------------------------------------------------------------------------------------------------
// k_connectivity: a member pair is a "root" when no face neighbour that is a member of the same solid has a smaller key.  Exactly one
// root  =>  the solid's vertex-inside cell set is face connected.
// ------------------------------------------------------------------------------------------------
struct ConnParams {
    DevMesh m;
    const DevSolid *solids;
    const unsigned char *n_item;
    const int *slots;
    int K;
    int *root_count; // [n_solids] zeroed
    const unsigned char *tile_proven;   // per tile: every binned candidate is provably connected
    const StepStatus *status;
};

// The key of a (cell, solid) pair is the fp32 squared distance between the fp32 copies of the cell centre and the solid
// centre, ties broken by the cell position: any strict total order certifies connectivity, this one is cheap and mostly
// decreases towards the solid centre.
__device__ __forceinline__ bool key_less(float ka, int ca, float kb, int cb) { return ka < kb || (ka == kb && ca < cb); }
__device__ __forceinline__ float conn_key(float4 p, const float *x) {
    const float dx = p.x - x[0], dy = p.y - x[1], dz = p.z - x[2];
    return dx * dx + dy * dy + dz * dz;
}

// slot index of solid s among the members of cell nb, or -1
__device__ __forceinline__ int find_member(const unsigned char *n_item, const int *slots, long long nC, int nb, int s) {
    const int nn = n_item[nb];
    for (int jj = 0; jj < nn; ++jj) {
        const int e = slots[(long long)jj * nC + nb];
        if ((e >> 3) == s) return (e & 3) ? jj : -1;
    }
    return -1;
}

// certificate of one cell: every member pair without a smaller-keyed member neighbour is a root of its solid
__device__ __forceinline__ void conn_cell(const ConnParams &P, int c) {
    const int n = P.n_item[c];
    if (n == 0) return;
    const long long nC = P.m.n_cells;
    const float4 p = __ldg(P.m.cc32 + c);
    for (int j = 0; j < n; ++j) {
        const int e = P.slots[(long long)j * nC + c];
        if ((e & 3) == 0) continue;
        const int s = e >> 3;
        const DevSolid &S = P.solids[s];
        if (S.conn_proven) continue;   // connected by construction (k_solid_prepare)
        const float *x = S.pos32;
        const float kc = conn_key(p, x);
        // try the face neighbour that lies towards the solid centre first: almost always a member with a smaller key
        const float rx = p.x - x[0], ry = p.y - x[1], rz = p.z - x[2];
        const float ax = fabsf(rx), ay = fabsf(ry), az = fabsf(rz);
        const int axis = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
        const float comp = axis == 0 ? rx : axis == 1 ? ry : rz;
        const int first = __ldg(P.m.nb6 + 6 * (long long)c + 2 * axis + (comp > 0 ? 0 : 1));
        bool has_parent = false;
        if (first >= 0) {
            const float4 pn = __ldg(P.m.cc32 + first);
            const float kn = conn_key(pn, x);
            if (key_less(kn, first, kc, c)) {
                // k_classify's own fp32 test: a neighbour certainly inside the solid is a member (ALL_INSIDE) without looking it up
                const float ri = S.ri32 - pn.w;
                has_parent = (ri > 0.f && kn < ri * ri) || find_member(P.n_item, P.slots, nC, first, s) >= 0;
            }
        }
        if (!has_parent) {
            const int nb0 = __ldg(P.m.nb_off + c), nb1 = __ldg(P.m.nb_off + c + 1);
            for (int k = nb0; k < nb1 && !has_parent; ++k) {
                const int nb = __ldg(P.m.nb + k);
                if (nb == first) continue;
                if (find_member(P.n_item, P.slots, nC, nb, s) >= 0 && key_less(conn_key(__ldg(P.m.cc32 + nb), x), nb, kc, c)) has_parent = true;
            }
        }
        if (!has_parent) atomicAdd(P.root_count + s, 1);
    }
}


__global__ void __launch_bounds__(256) k_connectivity(ConnParams P) {
    // nothing to certify: every binned solid is connected by construction and no plane / tilted 2-D solid is about
    if (!P.status->need_cert && P.status->n_global == 0) return;
    const long long nC = P.m.n_cells;
    const bool global = P.status->n_global != 0;
    // persistent; FOUR position blocks per iteration: the tile_key -> tile_proven chain of the four is in flight at once (with one
    // block per iteration the pass was a serial chain of ~55 two-level round trips per thread at 256^3 cells)
    for (long long c0 = (long long)blockIdx.x * 1024; c0 < nC; c0 += (long long)gridDim.x * 1024) {
        int c[4];
        bool skip[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long cq = c0 + q * 256 + threadIdx.x;
            c[q] = cq < nC ? (int)cq : -1;
        }
        unsigned t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = c[q] >= 0 ? __ldg(P.m.tile_key + c[q]) : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) skip[q] = c[q] < 0 || (!global && P.tile_proven[t[q]]);   // every candidate of the cell's tile is connected by construction
#pragma unroll
        for (int q = 0; q < 4; ++q) if (!skip[q]) conn_cell(P, c[q]);
    }
}

__global__ void k_snapshot(const unsigned long long *a, const unsigned long long *b, unsigned long long *dst) { dst[0] = *a; dst[1] = *b; }

// per-step totals for the status word + rhof scaling of the per-solid sums (solidcloud.cpp:424-425)
// rhof scaling alone (solidcloud.cpp:424-425): the split multi-GPU step scales right behind k_final, so that the all-reduce of the
// per-solid sums can run on its own stream while the certificate pass and the status totals follow on the context stream
__global__ void k_scale_ft(double *ft, long long n, const double *scal) {
    const double rhof = __ldg(scal + 1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) ft[i] = ft[i] * rhof;
}

__global__ void k_finalize(const unsigned *pair_counts, const int *root_count, int n_solids, StepStatus *status, double *ft, const double *scal, int scale) {
    unsigned long long c0 = 0, c1 = 0, c2 = 0;
    int nf = 0;
    const double rhof = __ldg(scal + 1);
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_solids; s += gridDim.x * blockDim.x) {
        c0 += pair_counts[3 * s];
        c1 += pair_counts[3 * s + 1];
        c2 += pair_counts[3 * s + 2];
        nf += root_count[s] > 1;
        if (scale) {
#pragma unroll
            for (int k = 0; k < 6; ++k) ft[6 * (long long)s + k] = ft[6 * (long long)s + k] * rhof;
        }
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

// ------------------------------------------------------------------------------------------------
// exact flood-fill replay for solids that failed the certificate (rare path, all on the GPU)
// ------------------------------------------------------------------------------------------------
struct ReplayParams {
    DevMesh m;
    const DevSolid *solids;
    const unsigned char *n_item;
    const int *slots;
    int K;
    const int *root_count;
    int *labels;                    // [n_cells*K] component label (min caller cell label) of flagged pairs
    const int *inv;                 // caller's cell label -> position
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
    const int n = P.n_item[c];
        for (int j = 0; j < n; ++j) {
        const int e = P.slots[(long long)j * P.m.n_cells + c];
        if ((e & 3) != 0 && P.root_count[e >> 3] > 1) P.labels[(long long)c * P.K + j] = __ldg(P.m.orig + c);
    }
}

// One sweep of the min-label propagation.  The host enqueues REPLAY_BATCH sweeps per round trip; `changed[sweep]` records whether
// sweep `sweep` of the batch lowered a label, and a sweep whose predecessor changed nothing returns at once (converged), so the
// convergence test runs on the device and the host reads ONE flag per batch.
#define REPLAY_BATCH 32
__global__ void k_replay_propagate(ReplayParams P, int sweep) {
    if (sweep > 0 && ((volatile int *)P.changed)[sweep - 1] == 0) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.m.n_cells) return;
    const int n = P.n_item[c];
    if (n == 0) return;
        const int nb0 = __ldg(P.m.nb_off + c), nb1 = __ldg(P.m.nb_off + c + 1);
        for (int j = 0; j < n; ++j) {
        int lab = P.labels[(long long)c * P.K + j];
        if (lab < 0) continue;
        const int s = P.slots[(long long)j * P.m.n_cells + c] >> 3;
        int best = lab;
        for (int k = nb0; k < nb1; ++k) {
            const int nb = __ldg(P.m.nb + k);
            const int jj = find_member(P.n_item, P.slots, P.m.n_cells, nb, s);
            if (jj >= 0) {
                const int l2 = ((volatile int *)P.labels)[(long long)nb * P.K + jj];
                if (l2 >= 0 && l2 < best) best = l2;
            }
        }
        if (best < lab) { P.labels[(long long)c * P.K + j] = best; P.changed[sweep] = 1; }
    }
}

// the solids that failed the certificate, as a compact list
__global__ void k_replay_flagged(ReplayParams P, int *flagged, int *n_flagged) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.n_solids && P.root_count[s] > 1) flagged[atomicAdd(n_flagged, 1)] = s;
}

// The reference seeds its flood fill at meshSearch::findNearestCell(centre) (src/solidcloud.cpp:363-365): the cell whose CENTRE is
// nearest the solid's centre over the WHOLE mesh — also when the solid's centre lies outside the (sub)mesh, where the nearest
// cell need not be a candidate of the solid at all (the fill then starts from the first member cell in index order,
// src/cellenumerator.cpp:52-63).  So every cell is compared with every flagged solid: a full pass, on the rare path only.
__global__ void k_replay_seed(ReplayParams P, const int *flagged, const int *n_flagged, int pass) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.m.n_cells) return;
    const D3 cc = ld3(P.m.cc, c);
    const int nf = *n_flagged;
    for (int t = 0; t < nf; ++t) {
        const int s = flagged[t];
        const D3 x = {P.solids[s].pos[0], P.solids[s].pos[1], P.solids[s].pos[2]};
        const unsigned long long key = (unsigned long long)__double_as_longlong(magSqr3(cc - x));
        if (pass == 0) atomicMin(P.seed_key + s, key);
        else if (key == P.seed_key[s]) atomicMin(P.seed_cell + s, __ldg(P.m.orig + c));
    }
    if (pass == 1) {
        const int n = P.n_item[c];
        for (int j = 0; j < n; ++j) {
            const int lab = P.labels[(long long)c * P.K + j];
            if (lab >= 0) atomicMin(P.min_label + (P.slots[(long long)j * P.m.n_cells + c] >> 3), lab);
        }
    }
}

__global__ void k_replay_choose(ReplayParams P) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_solids) return;
    if (P.root_count[s] <= 1) { P.chosen[s] = -1; return; }
    int chosen = P.min_label[s]; // component of the first member cell in index order (cellenumerator.cpp:52-63)
    const int go = P.seed_cell[s];   // caller's label of the nearest cell centre (ties: lowest label)
    if (go >= 0 && go < P.m.n_cells) {
        const int g = P.inv[go];
        const int jj = find_member(P.n_item, P.slots, P.m.n_cells, g, s);
        if (jj >= 0) chosen = P.labels[(long long)g * P.K + jj]; // the nearest cell is a member: it is the seed
    }
    P.chosen[s] = chosen;
}

__global__ void k_replay_mark(ReplayParams P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.m.n_cells) return;
    const int n = P.n_item[c];
        for (int j = 0; j < n; ++j) {
        const int lab = P.labels[(long long)c * P.K + j];
        if (lab >= 0 && lab != P.chosen[P.slots[(long long)j * P.m.n_cells + c] >> 3]) P.excluded[(long long)c * P.K + j] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// fixInternal (solidcloud.cpp:288-301)
// ------------------------------------------------------------------------------------------------
// Two consecutive cells per thread: Ct streams in as 16-byte loads (8 nC of the pass's 8 nC + 48 #(Ct >= 4) bytes).
__device__ __forceinline__ void fix_internal_cell(const double *cc, const sdfibm_solid_t *solids, int n_solids, double ct, double *U, int c) {
    if (ct >= 4) {
        const int id = (int)(ct - 4);
        if (id < n_solids) {
            const sdfibm_solid_t &S = solids[id];
            const D3 x = {S.pos[0], S.pos[1], S.pos[2]};
            const D3 u = D3{S.vel[0], S.vel[1], S.vel[2]} + cross3(D3{S.omega[0], S.omega[1], S.omega[2]}, ld3(cc, c) - x);
            U[3 * (long long)c] = u.x;
            U[3 * (long long)c + 1] = u.y;
            U[3 * (long long)c + 2] = u.z;
        }
    }
}
__global__ void __launch_bounds__(256) k_fix_internal(const double *__restrict__ cc, const sdfibm_solid_t *__restrict__ solids, int n_solids,
                                                      const double *__restrict__ Ct, double *__restrict__ U, int c_begin, int c_end) {
    const int c = c_begin + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= c_end) return;
    if ((c_begin & 1) == 0 && c + 1 < c_end) {
        const double2 ct = __ldg(reinterpret_cast<const double2 *>(Ct + c));
        fix_internal_cell(cc, solids, n_solids, ct.x, U, c);
        fix_internal_cell(cc, solids, n_solids, ct.y, U, c + 1);
    } else {
        fix_internal_cell(cc, solids, n_solids, Ct[c], U, c);
        if (c + 1 < c_end) fix_internal_cell(cc, solids, n_solids, Ct[c + 1], U, c + 1);
    }
}

// ------------------------------------------------------------------------------------------------
// the step either side of interact (src/main.cpp:70-77), on the device: U = U - Fs dt, T = (1 - As) T + Ts.  Only cells that hold a
// slot record can have non-zero Fs / As / Ts; everywhere else both updates are the identity (U - 0 dt, 1 T + 0), bit for bit.
// ------------------------------------------------------------------------------------------------
__global__ void k_apply_forcing(const unsigned char *__restrict__ n_item, const int *__restrict__ orig, int n_cells, const double *__restrict__ As,
                                const double *__restrict__ Fs, const double *__restrict__ Ts, double dt, double *__restrict__ U, double *__restrict__ T) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells || n_item[c] == 0) return;
    const long long oc = __ldg(orig + c);
    // every load first (one memory latency, not one per component), then the stores
    double u0 = 0, u1 = 0, u2 = 0, f0 = 0, f1 = 0, f2 = 0, t = 0, as = 0, ts = 0;
    if (U) {
        f0 = __ldg(Fs + 3 * oc); f1 = __ldg(Fs + 3 * oc + 1); f2 = __ldg(Fs + 3 * oc + 2);
        u0 = U[3 * oc]; u1 = U[3 * oc + 1]; u2 = U[3 * oc + 2];
    }
    if (T) { as = __ldg(As + oc); ts = __ldg(Ts + oc); t = T[oc]; }
    if (U) {
        U[3 * oc] = u0 - f0 * dt;
        U[3 * oc + 1] = u1 - f1 * dt;
        U[3 * oc + 2] = u2 - f2 * dt;
    }
    if (T) T[oc] = (1.0 - as) * t + ts;
}

// compact records of the touched cells (cells that hold a slot record): flags -> exclusive scan -> gather
__global__ void k_touched_flags(const unsigned char *n_item, int n_cells, int *flag) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= n_cells) flag[c] = (c < n_cells && n_item[c] != 0) ? 1 : 0;
}
__global__ void k_touched_gather(const unsigned char *n_item, const int *orig, const int *off, int n_cells, long long cap, const double *As,
                                 const double *Fs, const double *Ts, const double *Ct, int *cells, double *oAs, double *oFs, double *oTs, double *oCt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells || n_item[c] == 0) return;
    const long long o = off[c];
    if (o >= cap) return;
    const long long oc = __ldg(orig + c);
    cells[o] = (int)oc;
    oAs[o] = As[oc];
    oFs[3 * o] = Fs[3 * oc]; oFs[3 * o + 1] = Fs[3 * oc + 1]; oFs[3 * o + 2] = Fs[3 * oc + 2];
    oTs[o] = Ts[oc];
    oCt[o] = Ct[oc];
}

// ------------------------------------------------------------------------------------------------
// candidate list extraction (parity output, off the timed path): pairs in position order, a stable radix sort by the
// caller's cell label, then a stable radix sort by (solid, type): ascending cell ids inside every segment (std::set order).
// ------------------------------------------------------------------------------------------------
__global__ void k_list_count(const unsigned char *n_item, const int *slots, const unsigned char *excluded, int K, int n_cells, int *cnt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    int n = 0;
    const int ni = n_item[c];
    for (int j = 0; j < ni; ++j)
        if ((slots[(long long)j * n_cells + c] & 3) != 0 && !(excluded && excluded[(long long)c * K + j])) ++n;
    cnt[c] = n;
}
__global__ void k_list_emit(const unsigned char *n_item, const int *slots, const unsigned char *excluded, int K, int n_cells,
                            const int *off, const int *orig, unsigned *keys, int *vals) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    int o = off[c];
    const int ni = n_item[c];
    for (int j = 0; j < ni; ++j) {
        const int e = slots[(long long)j * n_cells + c];
        if ((e & 3) == 0 || (excluded && excluded[(long long)c * K + j])) continue;
        keys[o] = (unsigned)(3 * (e >> 3) + ((e & 3) - 1));
        vals[o] = orig[c];   // caller's cell label
        ++o;
    }
}
