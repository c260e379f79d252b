// sdfshape.h — composed shapes: what a plugin written from the reference's src/libshape/template.h does by hand in its isInside /
// signedDistance pair (sdf.h's primitives on transformed points, combined by sdf::U / I / D), stated ONCE as a post-fix program
// (include/sdfibm_b200.h, sdfibm_sdf_op_t) that the device evaluates as it is: a new composed shape is a constructor, not a
// change to the CUDA switch.
//
//   class Capsule2D : public SdfShape, public _shapecreator<Capsule2D> {
//   public:
//       Capsule2D(const dictionary &para) {
//           const scalar r = Foam::readScalar(para.lookup("radius")), l = Foam::readScalar(para.lookup("length"));
//           m_program.point2d().offset(vector(-l, 0, 0)).circle(r)          // sdf::circle(sdf::offset(p2d, (-l,0,0)), r)
//                    .point2d().offset(vector(l, 0, 0)).circle(r).unite()   // U({.., sdf::circle(sdf::offset(p2d, (l,0,0)), r)})
//                    .point2d().rectangle(l, r).unite();                    // U({.., sdf::rectangle(p2d, l, r)})
//           setBounds(/*outer*/ l + r, /*inner*/ r, /*two_d*/ true);
//           ... mass properties as in template.h ...
//       }
//       SHAPETYPENAME("Capsule2D")
//       virtual std::string description() const override { return "capsule"; }
//   };
//   REGISTERSHAPE(Capsule2D);
#pragma once
#include <vector>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
#endif
#include "../../csrc/device_math.cuh"
#include "ishape.h"

namespace sdfibm {

class SdfProgram {
    std::vector<sdfibm_sdf_op_t> m_ops;
    SdfProgram &add(int op, double a0 = 0, double a1 = 0, double a2 = 0) {
        sdfibm_sdf_op_t o{};
        o.op = op;
        o.a[0] = a0; o.a[1] = a1; o.a[2] = a2;
        m_ops.push_back(o);
        return *this;
    }

public:
    const std::vector<sdfibm_sdf_op_t> &ops() const { return m_ops; }
    SdfProgram &point() { return add(SDFIBM_OP_POINT); }          // com + p
    SdfProgram &point2d() { return add(SDFIBM_OP_POINT_2D); }     // ... with z = 0
    SdfProgram &offset(const vector &v) { return add(SDFIBM_OP_OFFSET, v.x(), v.y(), v.z()); }
    SdfProgram &rot30() { return add(SDFIBM_OP_ROT30); }
    SdfProgram &rot45() { return add(SDFIBM_OP_ROT45); }
    SdfProgram &rot60() { return add(SDFIBM_OP_ROT60); }
    SdfProgram &rot90() { return add(SDFIBM_OP_ROT90); }
    SdfProgram &rotth(scalar th) { return add(SDFIBM_OP_ROTTH, th); }
    SdfProgram &flipx() { return add(SDFIBM_OP_FLIPX); }
    SdfProgram &flipy() { return add(SDFIBM_OP_FLIPY); }
    SdfProgram &circle(scalar r) { return add(SDFIBM_OP_CIRCLE, r, r * r); }
    SdfProgram &sphere(scalar r) { return circle(r); }
    SdfProgram &rectangle(scalar ra, scalar rb) { return add(SDFIBM_OP_RECTANGLE, ra, rb); }
    SdfProgram &box(scalar ra, scalar rb, scalar rc) { return add(SDFIBM_OP_BOX, ra, rb, rc); }
    SdfProgram &ellipse(scalar a, scalar b) { return add(SDFIBM_OP_ELLIPSE, 1.0 / (a * a), 1.0 / (b * b)); }
    SdfProgram &ellipsoid(scalar a, scalar b, scalar c) { return add(SDFIBM_OP_ELLIPSOID, 1.0 / (a * a), 1.0 / (b * b), 1.0 / (c * c)); }
    SdfProgram &halfspace() { return add(SDFIBM_OP_HALFSPACE); }
    SdfProgram &unite() { return add(SDFIBM_OP_UNION); }
    SdfProgram &intersect() { return add(SDFIBM_OP_INTERSECT); }
    SdfProgram &subtract() { return add(SDFIBM_OP_DIFF); }
};

class SdfShape : public IShape {
protected:
    SdfProgram m_program;
    scalar m_rOut{0}, m_rIn{0};
    bool m_twoD{false};
    // certified radii about the body origin (before `com` is added): no point farther than `outer` is inside, every point closer
    // than `inner` is (0 = unknown); two_d: the shape ignores the body z coordinate
    void setBounds(scalar outer, scalar inner, bool two_d) { m_rOut = outer; m_rIn = inner; m_twoD = two_d; }

public:
    virtual bool lowerProgram(sdfibm_shape_t &out, std::vector<sdfibm_sdf_op_t> &table) const override {
        lowerCommon(out, SDFIBM_SHAPE_PROGRAM);
        out.p[0] = (double)table.size();
        out.p[1] = (double)m_program.ops().size();
        out.p[2] = m_rOut;
        out.p[3] = m_rIn;
        out.p[4] = m_twoD ? 1.0 : 0.0;
        table.insert(table.end(), m_program.ops().begin(), m_program.ops().end());
        return true;
    }

private:
    // host-side evaluation: the same interpreter the kernels run (device_math.cuh compiled for the host)
    virtual bool isInside(const vector &p) const override {
        double phi;
        return sdf_program_eval<false>(m_program.ops().data(), (int)m_program.ops().size(), D3{m_com.x(), m_com.y(), m_com.z()}, D3{p.x(), p.y(), p.z()}, phi);
    }
    virtual scalar signedDistance(const vector &p) const override {
        double phi;
        sdf_program_eval<true>(m_program.ops().data(), (int)m_program.ops().size(), D3{m_com.x(), m_com.y(), m_com.z()}, D3{p.x(), p.y(), p.z()}, phi);
        return phi;
    }
};

} // namespace sdfibm
