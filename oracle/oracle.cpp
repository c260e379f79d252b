// oracle.cpp — CPU restatement of the sdfibm hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (sdfibm_b200/) never links, imports or calls it.
//
// It restates, in plain fp64 C++ without OpenFOAM (build: -O2 -ffp-contract=off), the algorithm of
//   SolidCloud::interact / solidFluidInteract / checkAlpha / fixInternal   reference src/solidcloud.cpp:288-301,361-464,564-570
//   CellEnumerator (flood fill)                                           reference src/cellenumerator.cpp:6-78
//   GeometricTools (apex / pyramid volume fraction)                       reference src/geometrictools.cpp:6-116
//   IShape::world2local + the nine shape SDFs                              reference src/libshape/ishape.h:43-52, sdf/sdf.h, *.h
//   UGrid broad phase + narrow phase + force law                          reference src/libcollision/ugrid.{h,cpp}, collision.cpp, solidcloud.cpp:477-519
// OpenFOAM 12 (not vendored in the reference; only pin is README.md:9) supplies vector / quaternion
// arithmetic; its published formulas are restated in vec/quat helpers below.
//
// Parity pin: G1 (tool_vof/example/0/alpha.water) and G2 (examples/flow_past_cylinder/re200/0/As),
// see tests/test_oracle_golden.py.  Everything the goldens do not cover (Fs, Ts, Ct, force/torque,
// 3-D shapes, collisions) is pinned oracle-vs-GPU only ("parity unpinned" by the reference itself).
#include "../include/sdfibm_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// OpenFOAM vector / quaternion arithmetic (VectorI.H, quaternionI.H), operation order preserved.
// ---------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }          // operator&
inline V3 cross(V3 a, V3 b) {                                                        // operator^
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double magSqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double mag(V3 a) { return std::sqrt(magSqr(a)); }

struct Q4 {
    double w;
    V3 v;
};
inline Q4 conjugate(Q4 q) { return {q.w, -q.v}; }
inline Q4 qmul(Q4 a, Q4 b) {
    return {a.w * b.w - dot(a.v, b.v), a.w * b.v + b.w * a.v + cross(a.v, b.v)};
}
// quaternion::transform(u) = (mulq0v(u) * conjugate(*this)).v(),  mulq0v(u) = (-(v & u), w*u + (v ^ u))
inline V3 qtransform(Q4 q, V3 u) {
    Q4 m = {-dot(q.v, u), q.w * u + cross(q.v, u)};
    return qmul(m, conjugate(q)).v;
}

inline V3 ld3(const double *p, int64_t i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

// ---------------------------------------------------------------------------------------------
// sdf:: helpers (reference src/libshape/sdf/sdf.h)
// ---------------------------------------------------------------------------------------------
const double TOL = 1e-8;      // sdf.h:9
const double SMALL_ = 1e-6;   // src/types.h:19

inline double sdf_filter(double phi) { return (std::fabs(phi) < TOL) ? -TOL : phi; }           // sdf.h:147-150
inline bool circle_bool_fast(V3 p, double rSQR) { return magSqr(p) < rSQR; }                   // sdf.h:19-22
inline double circle_sd(V3 p, double r) { return mag(p) - r; }                                 // sdf.h:23-26
inline bool rectangle_bool(V3 p, double ra, double rb) {                                       // sdf.h:29-32
    return std::fabs(p.x) < ra && std::fabs(p.y) < rb;
}
inline double rectangle_sd(V3 p, double ra, double rb) {                                       // sdf.h:33-40
    double dx = std::fabs(p.x) - ra;
    double dy = std::fabs(p.y) - rb;
    double dxp = std::max(0.0, dx);
    double dyp = std::max(0.0, dy);
    return std::sqrt(dxp * dxp + dyp * dyp) + std::min(0.0, std::max(dx, dy));
}
inline bool box_bool(V3 p, double ra, double rb, double rc) {                                  // sdf.h:43-46
    return std::fabs(p.x) < ra && std::fabs(p.y) < rb && std::fabs(p.z) < rc;
}
inline double box_sd(V3 p, double ra, double rb, double rc) {                                  // sdf.h:47-56
    double dx = std::fabs(p.x) - ra;
    double dy = std::fabs(p.y) - rb;
    double dz = std::fabs(p.z) - rc;
    double dxp = std::max(0.0, dx);
    double dyp = std::max(0.0, dy);
    double dzp = std::max(0.0, dz);
    return std::sqrt(dxp * dxp + dyp * dyp + dzp * dzp) + std::min(0.0, std::max(dz, std::max(dx, dy)));
}
inline bool ellipse_bool_fast(V3 p, double ia, double ib) {                                    // sdf.h:59-62
    return p.x * p.x * ia + p.y * p.y * ib < 1.0;
}
inline double ellipse_sd(V3 p, double ia, double ib) {                                         // sdf.h:63-69
    double X = p.x * p.x * ia;
    double Y = p.y * p.y * ib;
    return 0.5 * (X + Y - 1.0) / (std::sqrt(X * ia + Y * ib));
}
inline bool ellipsoid_bool_fast(V3 p, double ia, double ib, double ic) {                       // sdf.h:72-75
    return p.x * p.x * ia + p.y * p.y * ib + p.z * p.z * ic < 1.0;
}
inline double ellipsoid_sd(V3 p, double ia, double ib, double ic) {                            // sdf.h:76-83
    double X = p.x * p.x * ia;
    double Y = p.y * p.y * ib;
    double Z = p.z * p.z * ic;
    return 0.5 * (X + Y + Z - 1.0) / (std::sqrt(X * ia + Y * ib + Z * ic));
}
inline V3 rot30(V3 p) {                                                                         // sdf.h:91-94
    return {0.866025404 * p.x + 0.5 * p.y, 0.866025404 * p.y - 0.5 * p.x, 0.0};
}
inline V3 flipy(V3 p) { return {p.x, -p.y, p.z}; }                                             // sdf.h:115-118
inline V3 offset(V3 p, V3 o) { return p - o; }                                                 // sdf.h:123-126

inline V3 rot45(V3 p) { return 0.707106781 * V3{p.x + p.y, -p.x + p.y, 0.0}; }                 // sdf.h:95-98
inline V3 rot60(V3 p) { return {0.866025404 * p.y + 0.5 * p.x, -0.866025404 * p.x + 0.5 * p.y, 0.0}; }   // sdf.h:99-102
inline V3 rot90(V3 p) { return {p.y, -p.x, 0.0}; }                                              // sdf.h:103-106
inline V3 rotth(V3 p, double th) {                                                              // sdf.h:107-112
    double s = std::sin(th);
    double c = std::cos(th);
    return {p.x * c + p.y * s, -p.x * s + p.y * c, 0.0};
}
inline V3 flipx(V3 p) { return {-p.x, p.y, p.z}; }                                             // sdf.h:119-122

// Composed shapes (SDFIBM_SHAPE_PROGRAM): what a plugin written from src/libshape/template.h does in its isInside /
// signedDistance pair — sdf.h's primitives on transformed points combined by sdf::U / I / D (:131-142) — restated as an
// evaluation of the same post-fix op list the product takes (include/sdfibm_b200.h, sdfibm_sdf_op_t), with the functions above.
static const sdfibm_sdf_op_t *g_ops = nullptr;   // set by oracle_set_programs (test infrastructure: one table per process)

struct SdfValue { bool b; double d; };
static SdfValue program_value(const sdfibm_shape_t &s, V3 p) {
    const V3 com = {s.com[0], s.com[1], s.com[2]};
    std::vector<V3> pts;
    std::vector<SdfValue> vals;
    const sdfibm_sdf_op_t *ops = g_ops + (int)s.p[0];
    for (int i = 0; i < (int)s.p[1]; ++i) {
        const sdfibm_sdf_op_t &o = ops[i];
        auto pop = [&]() { V3 q = pts.back(); pts.pop_back(); return q; };
        switch (o.op) {
        case SDFIBM_OP_POINT: pts.push_back(com + p); break;
        case SDFIBM_OP_POINT_2D: { V3 q = com + p; q.z = 0.0; pts.push_back(q); break; }
        case SDFIBM_OP_OFFSET: pts.back() = offset(pts.back(), V3{o.a[0], o.a[1], o.a[2]}); break;
        case SDFIBM_OP_ROT30: pts.back() = rot30(pts.back()); break;
        case SDFIBM_OP_ROT45: pts.back() = rot45(pts.back()); break;
        case SDFIBM_OP_ROT60: pts.back() = rot60(pts.back()); break;
        case SDFIBM_OP_ROT90: pts.back() = rot90(pts.back()); break;
        case SDFIBM_OP_ROTTH: pts.back() = rotth(pts.back(), o.a[0]); break;
        case SDFIBM_OP_FLIPX: pts.back() = flipx(pts.back()); break;
        case SDFIBM_OP_FLIPY: pts.back() = flipy(pts.back()); break;
        case SDFIBM_OP_CIRCLE: { V3 q = pop(); vals.push_back({circle_bool_fast(q, o.a[1]), circle_sd(q, o.a[0])}); break; }
        case SDFIBM_OP_RECTANGLE: { V3 q = pop(); vals.push_back({rectangle_bool(q, o.a[0], o.a[1]), rectangle_sd(q, o.a[0], o.a[1])}); break; }
        case SDFIBM_OP_BOX: { V3 q = pop(); vals.push_back({box_bool(q, o.a[0], o.a[1], o.a[2]), box_sd(q, o.a[0], o.a[1], o.a[2])}); break; }
        case SDFIBM_OP_ELLIPSE: { V3 q = pop(); vals.push_back({ellipse_bool_fast(q, o.a[0], o.a[1]), ellipse_sd(q, o.a[0], o.a[1])}); break; }
        case SDFIBM_OP_ELLIPSOID: { V3 q = pop(); vals.push_back({ellipsoid_bool_fast(q, o.a[0], o.a[1], o.a[2]), ellipsoid_sd(q, o.a[0], o.a[1], o.a[2])}); break; }
        case SDFIBM_OP_HALFSPACE: { V3 q = pop(); vals.push_back({q.y < 0, q.y}); break; }
        case SDFIBM_OP_UNION: { SdfValue b = vals.back(); vals.pop_back(); SdfValue &a = vals.back(); a = {std::max(a.b, b.b), std::min(a.d, b.d)}; break; }        // sdf::U
        case SDFIBM_OP_INTERSECT: { SdfValue b = vals.back(); vals.pop_back(); SdfValue &a = vals.back(); a = {std::min(a.b, b.b), std::max(a.d, b.d)}; break; }    // sdf::I
        case SDFIBM_OP_DIFF: { SdfValue b = vals.back(); vals.pop_back(); SdfValue &a = vals.back(); a = {a.b && (!b.b), std::max(a.d, -b.d)}; break; }             // sdf::D
        }
    }
    return vals.empty() ? SdfValue{false, 0.0} : vals[0];
}

// isInside in the body frame (private virtual of each IShape subclass)
bool shape_is_inside(const sdfibm_shape_t &s, V3 p) {
    const V3 com = {s.com[0], s.com[1], s.com[2]};
    switch (s.tag) {
    case SDFIBM_SHAPE_PROGRAM: return program_value(s, p).b;
    case SDFIBM_SHAPE_PLANE: return p.y < 0;                                                   // plane.h:21-24
    case SDFIBM_SHAPE_CIRCLE: return circle_bool_fast(com + V3{p.x, p.y, 0.0}, s.p[1]);        // circle.h:38-41
    case SDFIBM_SHAPE_SPHERE: return circle_bool_fast(com + p, s.p[1]);                        // sphere.h:37-40
    case SDFIBM_SHAPE_ELLIPSE: {                                                               // ellipse.h:41-45
        V3 p2 = com + p; p2.z = 0.0;
        return ellipse_bool_fast(p2, s.p[2], s.p[3]);
    }
    case SDFIBM_SHAPE_ELLIPSOID: return ellipsoid_bool_fast(p, s.p[3], s.p[4], s.p[5]);        // ellipsoid.h:43-46
    case SDFIBM_SHAPE_RECTANGLE: {                                                             // rectangle.h:41-48
        V3 p2 = com + p; p2.z = 0.0;
        return rectangle_bool(p2, s.p[0], s.p[1]);
    }
    case SDFIBM_SHAPE_BOX: return box_bool(com + p, s.p[0], s.p[1], s.p[2]);                   // box.h:40-47
    case SDFIBM_SHAPE_CIRCLE_TAIL: {                                                           // circle_tail.h:45-51
        V3 p2 = com + p; p2.z = 0.0;
        bool b1 = rectangle_bool(offset(p2, V3{s.p[2], 0.0, 0.0}), s.p[2], s.p[3]);
        bool b2 = circle_bool_fast(V3{p2.x, p2.y, 0.0}, s.p[1]);
        return std::max(b1, b2);
    }
    case SDFIBM_SHAPE_CIRCLE_TWOTAIL: {                                                        // circle_twotail.h:45-54
        V3 p2 = com + p; p2.z = 0.0;
        bool dc = circle_bool_fast(p2, s.p[1]);
        bool d1 = rectangle_bool(offset(rot30(p2), V3{s.p[2], 0, 0}), s.p[2], s.p[3]);
        bool d2 = rectangle_bool(offset(rot30(flipy(p2)), V3{s.p[2], 0, 0}), s.p[2], s.p[3]);
        return std::max(dc, std::max(d1, d2));
    }
    }
    return false;
}

// signedDistance in the body frame
double shape_signed_distance(const sdfibm_shape_t &s, V3 p) {
    const V3 com = {s.com[0], s.com[1], s.com[2]};
    switch (s.tag) {
    case SDFIBM_SHAPE_PROGRAM: return sdf_filter(program_value(s, p).d);                       // sdf::filter, sdf.h:147-150
    case SDFIBM_SHAPE_PLANE: return p.y;                                                       // plane.h:25-28 (unfiltered)
    case SDFIBM_SHAPE_CIRCLE: return sdf_filter(circle_sd(com + V3{p.x, p.y, 0.0}, s.p[0]));   // circle.h:42-45
    case SDFIBM_SHAPE_SPHERE: return sdf_filter(circle_sd(com + p, s.p[0]));                   // sphere.h:41-44
    case SDFIBM_SHAPE_ELLIPSE: {                                                               // ellipse.h:46-52
        V3 p2 = com + p; p2.z = 0.0;
        return sdf_filter(ellipse_sd(p2, s.p[2], s.p[3]));
    }
    case SDFIBM_SHAPE_ELLIPSOID: return sdf_filter(ellipsoid_sd(p, s.p[3], s.p[4], s.p[5]));   // ellipsoid.h:47-52
    case SDFIBM_SHAPE_RECTANGLE: {                                                             // rectangle.h:49-57
        V3 p2 = com + p; p2.z = 0.0;
        return sdf_filter(rectangle_sd(p2, s.p[0], s.p[1]));
    }
    case SDFIBM_SHAPE_BOX: return sdf_filter(box_sd(com + p, s.p[0], s.p[1], s.p[2]));         // box.h:48-56
    case SDFIBM_SHAPE_CIRCLE_TAIL: {                                                           // circle_tail.h:52-59
        V3 p2 = com + p; p2.z = 0.0;
        double d1 = circle_sd(p2, s.p[0]);
        double d2 = rectangle_sd(offset(p2, V3{s.p[2], 0.0, 0.0}), s.p[2], s.p[3]);
        return sdf_filter(std::min(d1, d2));
    }
    case SDFIBM_SHAPE_CIRCLE_TWOTAIL: {                                                        // circle_twotail.h:55-64
        V3 p2 = com + p; p2.z = 0.0;
        double dc = circle_sd(p2, s.p[0]);
        double d1 = rectangle_sd(offset(rot30(p2), V3{s.p[2], 0, 0}), s.p[2], s.p[3]);
        double d2 = rectangle_sd(offset(rot30(flipy(p2)), V3{s.p[2], 0, 0}), s.p[2], s.p[3]);
        return sdf_filter(std::min(dc, std::min(d1, d2)));
    }
    }
    return 0.0;
}

// Solid as the path sees it (src/solid.h:110-124)
struct SolidView {
    V3 center;
    Q4 q;
    V3 vel, omega;
    const sdfibm_shape_t *shape;
    int id;
    // IShape::world2local (ishape.h:43-46)
    V3 w2l(V3 p) const { return qtransform(conjugate(q), p - center); }
    bool phi01(V3 p) const { return shape_is_inside(*shape, w2l(p)); }
    double phi(V3 p) const { return shape_signed_distance(*shape, w2l(p)); }
    V3 evalPointVelocity(V3 p) const { return vel + cross(omega, p - center); }
};

SolidView make_view(const sdfibm_solid_t &s, const sdfibm_shape_t *shapes, int id) {
    SolidView v;
    v.center = {s.pos[0], s.pos[1], s.pos[2]};
    v.q = {s.quat[0], {s.quat[1], s.quat[2], s.quat[3]}};
    v.vel = {s.vel[0], s.vel[1], s.vel[2]};
    v.omega = {s.omega[0], s.omega[1], s.omega[2]};
    v.shape = &shapes[s.shape];
    v.id = id;
    return v;
}

// ---------------------------------------------------------------------------------------------
// Oracle context: mesh view + an exact nearest-cell-centre search (stands in for
// meshSearch::findNearestCell, solidcloud.cpp:217,363; ties -> lowest cell id).
// ---------------------------------------------------------------------------------------------
struct Oracle {
    sdfibm_mesh_t m;
    int twoD;
    // uniform grid over cell centres for the nearest search
    int gn[3];
    double glo[3], gh[3];
    std::vector<int> goff, gidx;
    std::vector<int> stamp; // re-usable visited array for the non-faithful mode
    int stampv = 0;

    void build_grid() {
        const int64_t n = m.n_cells;
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int64_t c = 0; c < n; ++c)
            for (int d = 0; d < 3; ++d) {
                lo[d] = std::min(lo[d], m.cell_centres[3 * c + d]);
                hi[d] = std::max(hi[d], m.cell_centres[3 * c + d]);
            }
        double ext[3], vol = 1;
        int nd = 0;
        double emax = 0;
        for (int d = 0; d < 3; ++d) { ext[d] = hi[d] - lo[d]; emax = std::max(emax, ext[d]); }
        for (int d = 0; d < 3; ++d) {
            if (ext[d] > 1e-9 * emax) { vol *= ext[d]; ++nd; }
            else ext[d] = 0.0; // degenerate direction (one-cell-thick meshes)
        }
        double h = nd ? std::pow(vol / std::max<double>(1.0, n / 2.0), 1.0 / nd) : 1.0;
        if (!(h > 0)) h = 1.0;
        for (int d = 0; d < 3; ++d) {
            glo[d] = lo[d];
            gh[d] = h;
            gn[d] = std::max(1, (int)std::floor(ext[d] / h) + 1);
        }
        const int64_t nb = (int64_t)gn[0] * gn[1] * gn[2];
        goff.assign(nb + 1, 0);
        std::vector<int> bin(n);
        for (int64_t c = 0; c < n; ++c) {
            int b = bin_of(ld3(m.cell_centres, c));
            bin[c] = b;
            ++goff[b + 1];
        }
        for (int64_t b = 0; b < nb; ++b) goff[b + 1] += goff[b];
        gidx.resize(n);
        std::vector<int> cur(goff.begin(), goff.end() - 1);
        for (int64_t c = 0; c < n; ++c) gidx[cur[bin[c]]++] = (int)c;
    }
    void bin_ijk(V3 p, int ijk[3]) const {
        double pp[3] = {p.x, p.y, p.z};
        for (int d = 0; d < 3; ++d) {
            double t = std::floor((pp[d] - glo[d]) / gh[d]);
            ijk[d] = (int)std::min<double>(gn[d] - 1, std::max<double>(0, t));
        }
    }
    int bin_of(V3 p) const {
        int ijk[3];
        bin_ijk(p, ijk);
        return (ijk[2] * gn[1] + ijk[1]) * gn[0] + ijk[0];
    }
    int nearest_cell(V3 p) const {
        if (m.n_cells == 0) return -1;
        int c0[3];
        bin_ijk(p, c0);
        double best = 1e300;
        int bestc = -1;
        const int maxr = std::max(gn[0], std::max(gn[1], gn[2]));
        for (int r = 0; r <= maxr; ++r) {
            // all bins on the Chebyshev shell of radius r around c0
            for (int k = c0[2] - r; k <= c0[2] + r; ++k) {
                if (k < 0 || k >= gn[2]) continue;
                for (int j = c0[1] - r; j <= c0[1] + r; ++j) {
                    if (j < 0 || j >= gn[1]) continue;
                    const bool edge_kj = (k == c0[2] - r || k == c0[2] + r || j == c0[1] - r || j == c0[1] + r);
                    for (int i = c0[0] - r; i <= c0[0] + r; ++i) {
                        if (i < 0 || i >= gn[0]) continue;
                        if (!edge_kj && !(i == c0[0] - r || i == c0[0] + r)) continue;
                        int b = (k * gn[1] + j) * gn[0] + i;
                        for (int t = goff[b]; t < goff[b + 1]; ++t) {
                            int c = gidx[t];
                            double d2 = magSqr(ld3(m.cell_centres, c) - p);
                            if (d2 < best || (d2 == best && c < bestc)) { best = d2; bestc = c; }
                        }
                    }
                }
            }
            if (bestc >= 0) {
                // any bin outside shell r is at least (r * h - offset inside the clamped bin) away
                double pp[3] = {p.x, p.y, p.z};
                double guard = 1e300;
                for (int d = 0; d < 3; ++d) {
                    double lo_edge = glo[d] + (c0[d] - r) * gh[d];
                    double hi_edge = glo[d] + (c0[d] + r + 1) * gh[d];
                    if (c0[d] - r > 0) guard = std::min(guard, pp[d] - lo_edge);
                    if (c0[d] + r < gn[d] - 1) guard = std::min(guard, hi_edge - pp[d]);
                }
                if (guard == 1e300) break;                 // whole grid visited
                if (guard > 0 && guard * guard > best) break;
            }
        }
        return bestc;
    }
};

// ---------------------------------------------------------------------------------------------
// CellEnumerator (reference src/cellenumerator.{h,cpp})
// ---------------------------------------------------------------------------------------------
typedef std::function<bool(const V3 &)> Predicate;

struct Enumerator {
    const sdfibm_mesh_t &m;
    std::vector<int> own_ct;   // faithful mode: fresh vector<CELL_TYPE>(nCells) per solid (cellenumerator.cpp:49)
    int *ct;                   // cell type array in use
    int stamp_base;            // non-faithful: types are stored as stamp_base + type
    Predicate pred;
    bool own_count = false;    // NOT the reference: compare with the neighbour's own vertex count (see next())
    std::queue<int> queue;
    std::set<size_t> sets[5];  // IntersectionSet (cellenumerator.h:25)

    int get(int c) const { int v = ct[c] - stamp_base; return (v >= 1 && v <= 4) ? v : 0; }
    void set(int c, int t) { ct[c] = stamp_base + t; }

    int count_vertex_inside(int c) const {                                   // cellenumerator.cpp:36-45
        int n = 0;
        for (int k = m.cell_points_off[c]; k < m.cell_points_off[c + 1]; ++k)
            n += pred(ld3(m.points, m.cell_points[k]));
        return n;
    }
    int nverts(int c) const { return m.cell_points_off[c + 1] - m.cell_points_off[c]; }

    Enumerator(const sdfibm_mesh_t &mesh, const Predicate &p, int seed, int *shared_ct, int stamp, bool own = false)
        : m(mesh), pred(p), own_count(own) {                                                   // cellenumerator.cpp:47-78
        if (shared_ct) { ct = shared_ct; stamp_base = stamp; }
        else { own_ct.assign(m.n_cells, 0); ct = own_ct.data(); stamp_base = 0; }
        if (seed < 0 || count_vertex_inside(seed) == 0) {
            seed = -1;
            for (int c = 0; c < m.n_cells; ++c)
                if (count_vertex_inside(c) > 0) { seed = c; break; }
        }
        if (seed >= 0) {
            queue.push(seed);
            int t;
            if (count_vertex_inside(seed) == nverts(seed)) t = SDFIBM_CELL_ALL_INSIDE;
            else if (pred(ld3(m.cell_centres, seed))) t = SDFIBM_CELL_CENTER_INSIDE;
            else t = SDFIBM_CELL_CENTER_OUTSIDE;
            set(seed, t);
            sets[t].insert(seed);
        }
    }
    void next() {                                                              // cellenumerator.cpp:6-34
        int icur = queue.front();
        for (int k = m.cell_cells_off[icur]; k < m.cell_cells_off[icur + 1]; ++k) {
            int inb = m.cell_cells[k];
            if (get(inb) != SDFIBM_CELL_UNVISITED) continue;
            int n_in = count_vertex_inside(inb);
            if (n_in == 0) {
                set(inb, SDFIBM_CELL_ALL_OUTSIDE);
                sets[SDFIBM_CELL_ALL_OUTSIDE].insert(inb);
                continue;
            }
            queue.push(inb);
            int t;
            // (sic) the reference compares with the vertex count of the CURRENT cell, :25 — on meshes that mix cell types the
            // outcome then depends on which neighbour discovered the cell first; own_count is the order-free variant
            if (n_in == nverts(own_count ? inb : icur)) t = SDFIBM_CELL_ALL_INSIDE;
            else if (pred(ld3(m.cell_centres, inb))) t = SDFIBM_CELL_CENTER_INSIDE;
            else t = SDFIBM_CELL_CENTER_OUTSIDE;
            set(inb, t);
            sets[t].insert(inb);
        }
    }
    void intersect() {                                                         // cellenumerator.h:47-52
        while (!queue.empty()) { next(); queue.pop(); }
    }
};

// ---------------------------------------------------------------------------------------------
// GeometricTools (reference src/geometrictools.cpp)
// ---------------------------------------------------------------------------------------------
struct GeoTools {
    const sdfibm_mesh_t &m;
    std::unordered_map<int, double> cache; // vertexInd -> phi (geometrictools.h:16,30)

    explicit GeoTools(const sdfibm_mesh_t &mesh) : m(mesh) {}

    double update_cache(int v, const SolidView &s) {                           // :6-11
        auto it = cache.find(v);
        if (it == cache.end()) return cache[v] = s.phi(ld3(m.points, v));
        return it->second;
    }
    static double line_fraction(double a, double b) {                          // :13-23
        if (a > 0 && b > 0) return 0;
        if (a <= 0 && b <= 0) return 1;
        if (a > 0) return -b / (a - b);
        return -a / (b - a);
    }
    V3 apex(const int32_t *ids, int n) {                                       // :25-45
        V3 A = ld3(m.points, ids[0]);
        double phiA = cache[ids[0]];
        V3 B = {0, 0, 0};
        double phiB = 0.0;
        for (int i = 1; i < n; ++i) {
            B = ld3(m.points, ids[i]);
            phiB = cache[ids[i]];
            if (phiA * phiB <= 0) break;
        }
        return A - std::fabs(phiA) / (SMALL_ + std::fabs(phiA) + std::fabs(phiB)) * (A - B);
    }
    double face_area(const int32_t *ids, int n) {                              // :74-96
        V3 ap = apex(ids, n);
        std::vector<double> phiarr(n);
        for (int i = 0; i < n; ++i) phiarr[i] = cache[ids[i]];
        double area = 0.0;
        for (int i = 0; i < n; ++i) {
            double phiO = phiarr[i];
            double phiA = phiarr[(i + 1) % n];
            V3 O = ld3(m.points, ids[i]);
            V3 A = ld3(m.points, ids[(i + 1) % n]);
            area += std::fabs(0.5 * mag(cross(A - O, ap - O))) * line_fraction(phiO, phiA);
        }
        return area;
    }
    double face_area_fraction(const int32_t *ids, int n, int f) {              // :98-116
        int sign_sum = 0;
        for (int i = 0; i < n; ++i) {
            if (cache[ids[i]] > 0) ++sign_sum;
            else --sign_sum;
        }
        if (sign_sum == n) return 0.0;
        if (sign_sum == -n) return 1.0;
        return face_area(ids, n) / mag(ld3(m.face_areas, f));
    }
    double cell_volume(int c, const SolidView &s, bool twoD) {                 // :47-72
        const int32_t *vids = m.cell_points + m.cell_points_off[c];
        const int nv = m.cell_points_off[c + 1] - m.cell_points_off[c];
        for (int i = 0; i < nv; ++i) update_cache(vids[i], s);
        V3 ap = apex(vids, nv);
        if (twoD) ap.z = 0.0;
        double volume = 0.0;
        for (int k = m.cell_faces_off[c]; k < m.cell_faces_off[c + 1]; ++k) {
            int f = m.cell_faces[k];
            std::vector<int32_t> face(m.face_points + m.face_points_off[f],
                                      m.face_points + m.face_points_off[f + 1]); // copy, as :66
            double eps_f = face_area_fraction(face.data(), (int)face.size(), f);
            volume += (1.0 / 3.0) * eps_f * std::fabs(dot(ap - ld3(m.face_centres, f), ld3(m.face_areas, f)));
        }
        return volume;
    }
};

} // namespace

// =================================================================================================
// C interface (ctypes)
// =================================================================================================
extern "C" {

void *oracle_create(const sdfibm_mesh_t *mesh, int twoD) {
    Oracle *o = new Oracle();
    o->m = *mesh;
    o->twoD = twoD;
    o->build_grid();
    return o;
}
void oracle_destroy(void *h) { delete (Oracle *)h; }
// the op table of the composed shapes (kept by the caller)
void oracle_set_programs(const sdfibm_sdf_op_t *ops) { g_ops = ops; }

int oracle_nearest_cell(void *h, const double p[3]) { return ((Oracle *)h)->nearest_cell({p[0], p[1], p[2]}); }

// phi01 / phi of one solid at n world points (shape unit tests)
void oracle_eval_points(const sdfibm_shape_t *shapes, const sdfibm_solid_t *solid, const double *pts, int64_t n,
                        int32_t *inside, double *phi) {
    SolidView s = make_view(*solid, shapes, 0);
    for (int64_t i = 0; i < n; ++i) {
        V3 p = ld3(pts, i);
        if (inside) inside[i] = s.phi01(p);
        if (phi) phi[i] = s.phi(p);
    }
}

/*
 * SolidCloud::interact (solidcloud.cpp:435-464).
 *   faithful != 0 keeps the reference's per-solid O(nCells) CELL_TYPE vector (cellenumerator.cpp:49);
 *   faithful == 0 reuses one stamped array (identical results, used for large parity cases).
 *   solid_begin/solid_end restrict the loop to a sub-range (bounded CPU baseline samples); fields
 *   then hold the contribution of that range only.
 *   list_off[3n+1]/list_cells (optional): the three sorted sets per solid, segment 3*s+type-1.
 *   timing_ms[0] = solid loop + checkAlpha (the reference's own timer, :442-451); [1] = whole call.
 * Returns 0, or 4 if list capacity is too small (list_off is still filled with sizes).
 */
int oracle_interact(void *h, const sdfibm_shape_t *shapes, const sdfibm_solid_t *solids, int n_solids,
                    int solid_begin, int solid_end, const double *U, double dt, double rhof, int faithful,
                    double *As, double *Fs, double *Ts, double *Ct, double *force_torque,
                    int32_t *list_off, int32_t *list_cells, int64_t list_cap, double *timing_ms) {
    using clk = std::chrono::high_resolution_clock;
    Oracle &o = *(Oracle *)h;
    const sdfibm_mesh_t &m = o.m;
    const int64_t nc = m.n_cells;
    auto t0 = clk::now();
    // :438-441 field resets
    std::fill(Ct, Ct + nc, 0.0);
    std::fill(As, As + nc, 0.0);
    std::fill(Fs, Fs + 3 * nc, 0.0);
    std::fill(Ts, Ts + nc, 0.0);
    if (force_torque) std::fill(force_torque, force_torque + 6 * (int64_t)n_solids, 0.0);
    const bool own_count = (faithful & 2) != 0;   // bit 1: order-free ALL_INSIDE test (not the reference's)
    faithful &= 1;
    if (!faithful && (int64_t)o.stamp.size() != nc) { o.stamp.assign(nc, 0); o.stampv = 0; }

    GeoTools geo(m);
    int64_t lpos = 0;
    int rc = 0;
    if (list_off) std::fill(list_off, list_off + 3 * (int64_t)n_solids + 1, 0);
    auto t1 = clk::now();
    for (int sid = solid_begin; sid < solid_end; ++sid) {
        SolidView solid = make_view(solids[sid], shapes, sid);
        // solidFluidInteract, :361-433
        int seed = o.nearest_cell(solid.center);
        Predicate pred = [&](const V3 &v) { return solid.phi01(v); };
        int *shared = nullptr;
        int stamp = 0;
        if (!faithful) {
            if (o.stampv > 2000000000) { std::fill(o.stamp.begin(), o.stamp.end(), 0); o.stampv = 0; }
            shared = o.stamp.data();
            stamp = o.stampv;
            o.stampv += 8;
        }
        Enumerator ce(m, pred, seed, shared, stamp, own_count);
        ce.intersect();
        const std::set<size_t> &sAI = ce.sets[SDFIBM_CELL_ALL_INSIDE];
        const std::set<size_t> &sCI = ce.sets[SDFIBM_CELL_CENTER_INSIDE];
        const std::set<size_t> &sCO = ce.sets[SDFIBM_CELL_CENTER_OUTSIDE];
        size_t num_inside = sAI.size();
        std::vector<size_t> cellids;
        cellids.reserve(sAI.size() + sCI.size() + sCO.size());
        cellids.insert(cellids.end(), sAI.begin(), sAI.end());
        cellids.insert(cellids.end(), sCI.begin(), sCI.end());
        cellids.insert(cellids.end(), sCO.begin(), sCO.end());
        if (list_off) {
            list_off[3 * sid + 1] = (int32_t)sAI.size();
            list_off[3 * sid + 2] = (int32_t)sCI.size();
            list_off[3 * sid + 3] = (int32_t)sCO.size();
            if (list_cells) {
                if (lpos + (int64_t)cellids.size() <= list_cap)
                    for (size_t k = 0; k < cellids.size(); ++k) list_cells[lpos + k] = (int32_t)cellids[k];
                else rc = 4;
                lpos += (int64_t)cellids.size();
            }
        }
        // :376-382
        int insideType = sid + 4;
        for (size_t c : sAI) Ct[c] = insideType;
        for (size_t c : sCI) Ct[c] = SDFIBM_CELL_CENTER_INSIDE;
        for (size_t c : sCO) Ct[c] = SDFIBM_CELL_CENTER_OUTSIDE;

        geo.cache.clear();                                                     // :395
        double dtINV = 1.0 / dt;
        V3 force = {0, 0, 0}, torque = {0, 0, 0};
        for (size_t counter = 0; counter < cellids.size(); ++counter) {        // :403-422
            size_t c = cellids[counter];
            double alpha = num_inside > 0 ? 1.0 : 0.0;
            if (counter >= num_inside) alpha = geo.cell_volume((int)c, solid, o.twoD != 0) / m.cell_volumes[c];
            As[c] += alpha;
            V3 cc = ld3(m.cell_centres, c);
            V3 uf = ld3(U, c);
            V3 us = solid.evalPointVelocity(cc);                               // :384-390
            V3 f_ = alpha * (uf - us);
            V3 t_ = cross(cc - solid.center, f_);
            force = force + f_ * m.cell_volumes[c] * dtINV;
            torque = torque + t_ * m.cell_volumes[c] * dtINV;
            V3 df = f_ * dtINV;
            Fs[3 * c] += df.x; Fs[3 * c + 1] += df.y; Fs[3 * c + 2] += df.z;
            Ts[c] += alpha;
        }
        force = force * rhof;                                                  // :424-425
        torque = torque * rhof;
        if (force_torque) {
            double *ft = force_torque + 6 * (int64_t)sid;
            ft[0] = force.x; ft[1] = force.y; ft[2] = force.z;
            ft[3] = torque.x; ft[4] = torque.y; ft[5] = torque.z;
        }
    }
    for (int64_t c = 0; c < nc; ++c) As[c] = std::min(As[c], 1.0);             // checkAlpha, :564-570
    auto t2 = clk::now();
    if (list_off) {                                                            // sizes -> offsets
        int32_t acc = 0;
        for (int64_t k = 1; k <= 3 * (int64_t)n_solids; ++k) { int32_t sz = list_off[k]; list_off[k] = acc + sz; acc += sz; }
        list_off[0] = 0;
    }
    if (timing_ms) {
        timing_ms[0] = std::chrono::duration<double, std::milli>(t2 - t1).count();
        timing_ms[1] = std::chrono::duration<double, std::milli>(t2 - t0).count();
    }
    return rc;
}

// SolidCloud::fixInternal (solidcloud.cpp:288-301)
void oracle_fix_internal(void *h, const sdfibm_shape_t *shapes, const sdfibm_solid_t *solids, int n_solids,
                         const double *Ct, double *U) {
    Oracle &o = *(Oracle *)h;
    for (int64_t c = 0; c < o.m.n_cells; ++c) {
        if (Ct[c] >= 4) {
            int id = (int)(Ct[c] - 4);
            if (id < 0 || id >= n_solids) continue;
            SolidView s = make_view(solids[id], shapes, id);
            V3 u = s.evalPointVelocity(ld3(o.m.cell_centres, c));
            U[3 * c] = u.x; U[3 * c + 1] = u.y; U[3 * c + 2] = u.z;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Collision step: UGrid (ugrid.h:29-57, ugrid.cpp:5-22,50-77), narrow phase (collision.cpp:7-57),
// force law (solidcloud.cpp:492-519).  delta < 0 reproduces HEAD (no pairs).
// ---------------------------------------------------------------------------------------------
static const char *kTypeName[SDFIBM_SHAPE_NTAGS] = {"Plane", "Circle", "Sphere", "Ellipse", "Ellipsoid",
                                                    "Rectangle", "Box", "Circle_Tail", "Circle_TwoTail"};

int oracle_collide(const double bmin[3], const double bmax[3], const sdfibm_shape_t *shapes,
                   const sdfibm_solid_t *solids, int n_solids, double delta, int32_t *pairs, int64_t cap,
                   int64_t *n_pairs, double *force_torque) {
    // UGrid ctor
    double deltaINV = 1.0 / delta;
    int nx = (int)std::ceil((bmax[0] - bmin[0]) * deltaINV);
    int ny = (int)std::ceil((bmax[1] - bmin[1]) * deltaINV);
    int nz = (int)std::ceil((bmax[2] - bmin[2]) * deltaINV);
    int nynz = ny * nz;
    long ncells = (long)nx * ny * nz;
    std::unordered_map<int, std::vector<int>> map;
    for (long i = 0; i < ncells; ++i) map[(int)i] = std::vector<int>();
    auto hash3 = [&](int i, int j, int k) { return i * nynz + j * nz + k; };
    // insert every solid centre (solidcloud.cpp:480-484)
    for (int s = 0; s < n_solids; ++s) {
        int i = (int)std::floor((solids[s].pos[0] - bmin[0]) * deltaINV);
        int j = (int)std::floor((solids[s].pos[1] - bmin[1]) * deltaINV);
        int k = (int)std::floor((solids[s].pos[2] - bmin[2]) * deltaINV);
        map[hash3(i, j, k)].push_back(s);
    }
    std::vector<std::pair<int, int>> cp;
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                int myid = hash3(i, j, k);
                if (map[myid].empty()) continue;
                for (int nbi = i - 1; nbi <= i + 1; ++nbi)
                    for (int nbj = j - 1; nbj <= j + 1; ++nbj)
                        for (int nbk = k - 1; nbk <= k + 1; ++nbk) {
                            if (nbi < 0 || nbi > nx - 1) continue;
                            if (nbj < 0 || nbj > ny - 1) continue;
                            if (nbk < 0 || nbk > nz - 1) continue;
                            int nbid = hash3(nbi, nbj, nbk);
                            if (map[nbid].empty()) continue;
                            std::vector<int> &lst = map[myid];
                            std::vector<int> &nbl = map[nbid];
                            for (int pi : lst)
                                for (int qi : nbl)
                                    if (pi < qi) cp.push_back({pi, qi});
                        }
            }
    *n_pairs = (int64_t)cp.size();
    int rc = 0;
    for (size_t t = 0; t < cp.size(); ++t) {
        if ((int64_t)t < cap && pairs) { pairs[2 * t] = cp[t].first; pairs[2 * t + 1] = cp[t].second; }
        else if (pairs) rc = 4;
    }
    if (!force_torque) return rc;
    // SHAPE2ID with std::map::operator[] semantics: unknown names map to 0 = "Plane" (collision.h:11-15, collision.cpp:54-57)
    std::map<std::string, int> SHAPE2ID = {{"Plane", 0}, {"Circle", 1}, {"Sphere", 2}};
    for (auto &pr : cp) {
        SolidView s1 = make_view(solids[pr.first], shapes, pr.first);
        SolidView s2 = make_view(solids[pr.second], shapes, pr.second);
        int a = SHAPE2ID[kTypeName[s1.shape->tag]];
        int b = SHAPE2ID[kTypeName[s2.shape->tag]];
        double cd;
        V3 cN;
        auto planeSphere = [&](const SolidView &p, const SolidView &s) {      // collision.cpp:22-30
            V3 sc = qtransform(conjugate(p.q), s.center - p.center);
            double p2s = sc.y;
            cN = qtransform(p.q, V3{0, 1, 0});
            cd = s.shape->radiusB - p2s;
        };
        if (a == 0 && b == 0) continue;                                        // table entry nullptr -> skip (:497-503)
        if (a == 0) planeSphere(s1, s2);
        else if (b == 0) planeSphere(s2, s1);                                  // spherePlane(s, p) = planeSphere(p, s)
        else if (a == b) {                                                     // sphereSphere / circleCircle (:7-18)
            V3 s2s = s2.center - s1.center;
            double ms = mag(s2s);
            cN = s2s / ms;                                                     // Foam::normalised
            cd = s1.shape->radiusB + s2.shape->radiusB - ms;
        } else continue;                                                       // Circle–Sphere: nullptr
        if (cd < 0) continue;                                                  // :509-510
        V3 force = (1e4 * cd) * cN;                                            // :511
        double *f1 = force_torque + 6 * (int64_t)pr.first;
        double *f2 = force_torque + 6 * (int64_t)pr.second;
        f1[0] += -force.x; f1[1] += -force.y; f1[2] += -force.z;               // :516-517 (torque is zero)
        f2[0] += force.x; f2[1] += force.y; f2[2] += force.z;
    }
    return rc;
}

} // extern "C"
