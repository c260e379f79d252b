// device_math.cuh — fp64 vector/quaternion arithmetic and the shape tagged union, device side.
//
// Operation ORDER mirrors OpenFOAM's VectorI.H / quaternionI.H as used by the reference
// (src/libshape/ishape.h:43-46) and the reference's sdf:: helpers (src/libshape/sdf/sdf.h), because
// the candidate lists depend on strict `<` predicates.  This translation unit is
// compiled with -fmad=false so no multiply-add is contracted.
#pragma once
#include "../../include/sdfibm_b200.h"

struct D3 {
    double x, y, z;
};
__host__ __device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ D3 operator-(D3 a) { return {-a.x, -a.y, -a.z}; }
__host__ __device__ __forceinline__ D3 operator*(double s, D3 a) { return {s * a.x, s * a.y, s * a.z}; }
__host__ __device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ D3 operator/(D3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
__host__ __device__ __forceinline__ double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ D3 cross3(D3 a, D3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__host__ __device__ __forceinline__ double magSqr3(D3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__host__ __device__ __forceinline__ double mag3(D3 a) { return sqrt(magSqr3(a)); }

struct DQ {
    double w;
    D3 v;
};
// quaternion::transform(u) = (mulq0v(u) * conjugate(q)).v()   (OpenFOAM quaternionI.H)
__host__ __device__ __forceinline__ D3 qtransform(DQ q, D3 u) {
    const double mw = -dot3(q.v, u);
    const D3 mv = q.w * u + cross3(q.v, u);
    const D3 cv = -q.v; // conjugate(q).v
    return mw * cv + q.w * mv + cross3(mv, cv);
}
// IShape::world2local: conjugate(q).transform(p - t)
__host__ __device__ __forceinline__ D3 world2local(DQ q, D3 t, D3 p) {
    DQ c = {q.w, -q.v};
    return qtransform(c, p - t);
}

// With the identity quaternion (w = 1, v = 0 — every solid that has not rotated yet) the sandwich above returns p - t
// exactly: each product with a zero component is a signed zero, and x + (+-0) == x.  Only the SIGN of an exactly-zero
// component can differ, which no predicate, distance or volume below can observe.  Callers test the quaternion once per
// warp and skip the 42 multiply / adds per point.
__host__ __device__ __forceinline__ bool quat_is_identity(DQ q) { return q.w == 1.0 && q.v.x == 0.0 && q.v.y == 0.0 && q.v.z == 0.0; }
__host__ __device__ __forceinline__ D3 world2local_sel(DQ q, D3 t, D3 p, bool identity) { return identity ? p - t : world2local(q, t, p); }

// std::max / std::min semantics ((a<b)?b:a, (b<a)?b:a)
__host__ __device__ __forceinline__ double smax(double a, double b) { return (a < b) ? b : a; }
__host__ __device__ __forceinline__ double smin(double a, double b) { return (b < a) ? b : a; }

#define SDF_TOL 1e-8
#define SDF_SMALL 1e-6

__host__ __device__ __forceinline__ double sdf_filter(double phi) { return (fabs(phi) < SDF_TOL) ? -SDF_TOL : phi; }
__host__ __device__ __forceinline__ bool rect_bool(D3 p, double ra, double rb) { return fabs(p.x) < ra && fabs(p.y) < rb; }
__host__ __device__ __forceinline__ double rect_sd(D3 p, double ra, double rb) {
    double dx = fabs(p.x) - ra;
    double dy = fabs(p.y) - rb;
    double dxp = smax(0.0, dx);
    double dyp = smax(0.0, dy);
    return sqrt(dxp * dxp + dyp * dyp) + smin(0.0, smax(dx, dy));
}
__host__ __device__ __forceinline__ D3 rot30(D3 p) { return {0.866025404 * p.x + 0.5 * p.y, 0.866025404 * p.y - 0.5 * p.x, 0.0}; }

// the other point transformations of sdf.h:95-122 (literals as in the reference)
__host__ __device__ __forceinline__ D3 rot45(D3 p) { return 0.707106781 * D3{p.x + p.y, -p.x + p.y, 0.0}; }
__host__ __device__ __forceinline__ D3 rot60(D3 p) { return {0.866025404 * p.y + 0.5 * p.x, -0.866025404 * p.x + 0.5 * p.y, 0.0}; }
__host__ __device__ __forceinline__ D3 rot90(D3 p) { return {p.y, -p.x, 0.0}; }
__host__ __device__ __forceinline__ D3 rotth(D3 p, double th) {
    const double s = sin(th), c = cos(th);
    return {p.x * c + p.y * s, -p.x * s + p.y * c, 0.0};
}

// A composed shape (SDFIBM_SHAPE_PROGRAM): post-fix evaluation of sdf.h's primitives / transformations / Boolean operations
// (see sdfibm_sdf_op_t).  The bool and the scalar of a value are the reference's two separate expressions, side by side.
// Programs are validated on the host (sdfibm_set_shapes): the stacks cannot over- or underflow here.
#ifdef __CUDACC__
#define SDF_NOINLINE __noinline__   // a real call, taken only by composed shapes: the interpreter's local stacks stay out of the hot kernels' frames
#else
#define SDF_NOINLINE
#endif
template <bool WANT_PHI>
__host__ __device__ SDF_NOINLINE bool sdf_program_eval(const sdfibm_sdf_op_t *ops, int n_ops, D3 com, D3 p, double &phi) {
    D3 pt[SDFIBM_SDF_STACK];
    bool vb[SDFIBM_SDF_STACK];
    double vd[SDFIBM_SDF_STACK];
    int np = 0, nv = 0;
    for (int i = 0; i < n_ops; ++i) {
        const sdfibm_sdf_op_t &o = ops[i];
        const double a0 = o.a[0], a1 = o.a[1], a2 = o.a[2];
        switch (o.op) {
        case SDFIBM_OP_POINT: pt[np++] = com + p; break;
        case SDFIBM_OP_POINT_2D: { D3 q = com + p; q.z = 0.0; pt[np++] = q; break; }
        case SDFIBM_OP_OFFSET: pt[np - 1] = pt[np - 1] - D3{a0, a1, a2}; break;
        case SDFIBM_OP_ROT30: pt[np - 1] = rot30(pt[np - 1]); break;
        case SDFIBM_OP_ROT45: pt[np - 1] = rot45(pt[np - 1]); break;
        case SDFIBM_OP_ROT60: pt[np - 1] = rot60(pt[np - 1]); break;
        case SDFIBM_OP_ROT90: pt[np - 1] = rot90(pt[np - 1]); break;
        case SDFIBM_OP_ROTTH: pt[np - 1] = rotth(pt[np - 1], a0); break;
        case SDFIBM_OP_FLIPX: pt[np - 1].x = -pt[np - 1].x; break;
        case SDFIBM_OP_FLIPY: pt[np - 1].y = -pt[np - 1].y; break;
        case SDFIBM_OP_CIRCLE: {
            const double m2 = magSqr3(pt[--np]);
            vb[nv] = m2 < a1;
            if (WANT_PHI) vd[nv] = sqrt(m2) - a0;
            ++nv;
            break;
        }
        case SDFIBM_OP_RECTANGLE: {
            const D3 P = pt[--np];
            vb[nv] = rect_bool(P, a0, a1);
            if (WANT_PHI) vd[nv] = rect_sd(P, a0, a1);
            ++nv;
            break;
        }
        case SDFIBM_OP_BOX: {
            const D3 P = pt[--np];
            const double dx = fabs(P.x) - a0, dy = fabs(P.y) - a1, dz = fabs(P.z) - a2;
            vb[nv] = fabs(P.x) < a0 && fabs(P.y) < a1 && fabs(P.z) < a2;
            if (WANT_PHI) {
                const double dxp = smax(0.0, dx), dyp = smax(0.0, dy), dzp = smax(0.0, dz);
                vd[nv] = sqrt(dxp * dxp + dyp * dyp + dzp * dzp) + smin(0.0, smax(dz, smax(dx, dy)));
            }
            ++nv;
            break;
        }
        case SDFIBM_OP_ELLIPSE: {
            const D3 P = pt[--np];
            const double X = P.x * P.x * a0, Y = P.y * P.y * a1;
            vb[nv] = X + Y < 1.0;
            if (WANT_PHI) vd[nv] = 0.5 * (X + Y - 1.0) / (sqrt(X * a0 + Y * a1));
            ++nv;
            break;
        }
        case SDFIBM_OP_ELLIPSOID: {
            const D3 P = pt[--np];
            const double X = P.x * P.x * a0, Y = P.y * P.y * a1, Z = P.z * P.z * a2;
            vb[nv] = X + Y + Z < 1.0;
            if (WANT_PHI) vd[nv] = 0.5 * (X + Y + Z - 1.0) / (sqrt(X * a0 + Y * a1 + Z * a2));
            ++nv;
            break;
        }
        case SDFIBM_OP_HALFSPACE: {
            const D3 P = pt[--np];
            vb[nv] = P.y < 0;
            if (WANT_PHI) vd[nv] = P.y;
            ++nv;
            break;
        }
        case SDFIBM_OP_UNION:       // sdf::U: std::min of the distances, std::max of the bools (:135-138)
            --nv;
            vb[nv - 1] = vb[nv - 1] || vb[nv];
            if (WANT_PHI) vd[nv - 1] = smin(vd[nv - 1], vd[nv]);
            break;
        case SDFIBM_OP_INTERSECT:   // sdf::I (:139-142)
            --nv;
            vb[nv - 1] = vb[nv - 1] && vb[nv];
            if (WANT_PHI) vd[nv - 1] = smax(vd[nv - 1], vd[nv]);
            break;
        case SDFIBM_OP_DIFF:        // sdf::D (:131-133)
            --nv;
            vb[nv - 1] = vb[nv - 1] && !vb[nv];
            if (WANT_PHI) vd[nv - 1] = smax(vd[nv - 1], -vd[nv]);
            break;
        default: break;
        }
    }
    if (WANT_PHI) phi = (nv > 0) ? sdf_filter(vd[0]) : 0.0;
    return nv > 0 && vb[0];
}

// Evaluate isInside and (optionally) signedDistance for a body-frame point.  Shape parameters are
// passed by pointer to the POD record so both host and device can call it.
// `ops`: the op table SDFIBM_SHAPE_PROGRAM records point into (may be null when the table holds no such record).
// HAS_PROG = false compiles the interpreter call out (the kernels of shape tables without composed shapes).
template <bool WANT_PHI, bool HAS_PROG = true>
__host__ __device__ __forceinline__ bool shape_eval(const sdfibm_shape_t &s, D3 p, double &phi, const sdfibm_sdf_op_t *ops = nullptr) {
    const D3 com = {s.com[0], s.com[1], s.com[2]};
    switch (s.tag) {
    case SDFIBM_SHAPE_PROGRAM:
        if (!HAS_PROG || !ops) break;
        return sdf_program_eval<WANT_PHI>(ops + (int)s.p[0], (int)s.p[1], com, p, phi);
    case SDFIBM_SHAPE_PLANE:
        if (WANT_PHI) phi = p.y;
        return p.y < 0;
    case SDFIBM_SHAPE_CIRCLE: {
        D3 P = com + D3{p.x, p.y, 0.0};
        double m2 = magSqr3(P);
        if (WANT_PHI) phi = sdf_filter(sqrt(m2) - s.p[0]);
        return m2 < s.p[1];
    }
    case SDFIBM_SHAPE_SPHERE: {
        D3 P = com + p;
        double m2 = magSqr3(P);
        if (WANT_PHI) phi = sdf_filter(sqrt(m2) - s.p[0]);
        return m2 < s.p[1];
    }
    case SDFIBM_SHAPE_ELLIPSE: {
        D3 P = com + p;
        double X = P.x * P.x * s.p[2];
        double Y = P.y * P.y * s.p[3];
        if (WANT_PHI) phi = sdf_filter(0.5 * (X + Y - 1.0) / (sqrt(X * s.p[2] + Y * s.p[3])));
        return X + Y < 1.0;
    }
    case SDFIBM_SHAPE_ELLIPSOID: {
        double X = p.x * p.x * s.p[3];
        double Y = p.y * p.y * s.p[4];
        double Z = p.z * p.z * s.p[5];
        if (WANT_PHI) phi = sdf_filter(0.5 * (X + Y + Z - 1.0) / (sqrt(X * s.p[3] + Y * s.p[4] + Z * s.p[5])));
        return X + Y + Z < 1.0;
    }
    case SDFIBM_SHAPE_RECTANGLE: {
        D3 P = com + p;
        P.z = 0.0;
        if (WANT_PHI) phi = sdf_filter(rect_sd(P, s.p[0], s.p[1]));
        return rect_bool(P, s.p[0], s.p[1]);
    }
    case SDFIBM_SHAPE_BOX: {
        D3 P = com + p;
        double dx = fabs(P.x) - s.p[0];
        double dy = fabs(P.y) - s.p[1];
        double dz = fabs(P.z) - s.p[2];
        if (WANT_PHI) {
            double dxp = smax(0.0, dx), dyp = smax(0.0, dy), dzp = smax(0.0, dz);
            phi = sdf_filter(sqrt(dxp * dxp + dyp * dyp + dzp * dzp) + smin(0.0, smax(dz, smax(dx, dy))));
        }
        return fabs(P.x) < s.p[0] && fabs(P.y) < s.p[1] && fabs(P.z) < s.p[2];
    }
    case SDFIBM_SHAPE_CIRCLE_TAIL: {
        D3 P = com + p;
        P.z = 0.0;
        D3 Po = P - D3{s.p[2], 0.0, 0.0};
        double m2 = magSqr3(P);
        if (WANT_PHI) phi = sdf_filter(smin(sqrt(m2) - s.p[0], rect_sd(Po, s.p[2], s.p[3])));
        return rect_bool(Po, s.p[2], s.p[3]) || (m2 < s.p[1]);
    }
    case SDFIBM_SHAPE_CIRCLE_TWOTAIL: {
        D3 P = com + p;
        P.z = 0.0;
        D3 P1 = rot30(P) - D3{s.p[2], 0.0, 0.0};
        D3 P2 = rot30(D3{P.x, -P.y, P.z}) - D3{s.p[2], 0.0, 0.0};
        double m2 = magSqr3(P);
        if (WANT_PHI) {
            double dc = sqrt(m2) - s.p[0];
            double d1 = rect_sd(P1, s.p[2], s.p[3]);
            double d2 = rect_sd(P2, s.p[2], s.p[3]);
            phi = sdf_filter(smin(dc, smin(d1, d2)));
        }
        return (m2 < s.p[1]) || rect_bool(P1, s.p[2], s.p[3]) || rect_bool(P2, s.p[2], s.p[3]);
    }
    }
    if (WANT_PHI) phi = 0.0;
    return false;
}
