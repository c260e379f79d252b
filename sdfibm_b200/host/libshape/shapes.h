// shapes.h — the nine shape plugins (reference src/libshape/{circle,sphere,ellipse,ellipsoid,rectangle,box,circle_tail,
// circle_twotail,plane}.h): dictionary keys, derived members, mass properties and bounding radius as the reference
// computes them.  isInside / signedDistance evaluate the SAME device_math.cuh code the kernels run (compiled for the
// host here), so there is one statement of the SDF arithmetic in the product.
#pragma once
#include <cmath>
#include <string>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
#endif
#include "../../csrc/device_math.cuh"
#include "ishape.h"
#include "sdfshape.h"

namespace sdfibm {

namespace detail {
inline bool eval_inside(const sdfibm_shape_t &r, const vector &p) {
    double phi;
    return shape_eval<false>(r, D3{p.x(), p.y(), p.z()}, phi);
}
inline scalar eval_sd(const sdfibm_shape_t &r, const vector &p) {
    double phi;
    shape_eval<true>(r, D3{p.x(), p.y(), p.z()}, phi);
    return phi;
}
inline void diag_moi(IShape &s, scalar a, scalar b, scalar c) {
    s.m_moi = tensor(a, 0, 0, 0, b, 0, 0, 0, c);
    s.m_moiINV = Foam::inv(s.m_moi);
}
} // namespace detail

#define SDFIBM_SHAPE_EVAL_VIA_RECORD()                                                                          \
    virtual bool isInside(const vector &p) const override { sdfibm_shape_t r; lower(r); return detail::eval_inside(r, p); } \
    virtual scalar signedDistance(const vector &p) const override { sdfibm_shape_t r; lower(r); return detail::eval_sd(r, p); }

class Circle : public IShape, public _shapecreator<Circle> {   // circle.h:15-46
    scalar m_radius, m_radiusSQR;

public:
    Circle(const dictionary &para) {
        m_radius = Foam::readScalar(para.lookup("radius"));
        m_com = para.lookupOrDefault("com", vector::zero);
        m_radiusSQR = m_radius * m_radius;
        m_volume = M_PI * m_radiusSQR;
        m_volumeINV = 1.0 / m_volume;
        const scalar tmp = 0.5 * m_volume * m_radiusSQR;
        detail::diag_moi(*this, tmp, tmp, tmp);
        m_radiusB = m_radius;
    }
    scalar getRadius() const { return m_radius; }
    scalar getVolume() const { return m_volume; }
    SHAPETYPENAME("Circle")
    virtual std::string description() const override { return "circle (x-y plane), r = " + std::to_string(m_radius); }
    virtual bool lower(sdfibm_shape_t &o) const override {
        lowerCommon(o, SDFIBM_SHAPE_CIRCLE);
        o.p[0] = m_radius;
        o.p[1] = m_radiusSQR;
        return true;
    }
    SDFIBM_SHAPE_EVAL_VIA_RECORD()
};

class Sphere : public IShape, public _shapecreator<Sphere> {   // sphere.h:14-45
    scalar m_radius, m_radiusSQR;

public:
    Sphere(const dictionary &para) {
        m_radius = Foam::readScalar(para.lookup("radius"));
        m_com = para.lookupOrDefault("com", vector::zero);
        m_radiusSQR = m_radius * m_radius;
        m_volume = 4.0 / 3.0 * M_PI * m_radiusSQR * m_radius;
        m_volumeINV = 1.0 / m_volume;
        const scalar tmp = 0.4 * m_volume * m_radiusSQR;
        detail::diag_moi(*this, tmp, tmp, tmp);
        m_radiusB = m_radius;
    }
    scalar getRadius() const { return m_radius; }
    SHAPETYPENAME("Sphere")
    virtual std::string description() const override { return "sphere, r = " + std::to_string(m_radius); }
    virtual bool lower(sdfibm_shape_t &o) const override {
        lowerCommon(o, SDFIBM_SHAPE_SPHERE);
        o.p[0] = m_radius;
        o.p[1] = m_radiusSQR;
        return true;
    }
    SDFIBM_SHAPE_EVAL_VIA_RECORD()
};

class Ellipse : public IShape, public _shapecreator<Ellipse> {   // ellipse.h:14-54
    scalar m_radiusa, m_radiusb, m_radiusaSQRINV, m_radiusbSQRINV;

public:
    Ellipse(const dictionary &para) {
        m_radiusa = Foam::readScalar(para.lookup("radiusa"));
        m_radiusb = Foam::readScalar(para.lookup("radiusb"));
        m_com = para.lookupOrDefault("com", vector::zero);
        m_radiusaSQRINV = 1.0 / (m_radiusa * m_radiusa);
        m_radiusbSQRINV = 1.0 / (m_radiusb * m_radiusb);
        m_volume = M_PI * m_radiusa * m_radiusb;
        m_volumeINV = 1.0 / m_volume;
        const scalar tmp = 0.25 * m_volume * (m_radiusa * m_radiusa + m_radiusb * m_radiusb);
        detail::diag_moi(*this, tmp, tmp, tmp);
        m_radiusB = std::max(m_radiusa, m_radiusb);
    }
    SHAPETYPENAME("Ellipse")
    virtual std::string description() const override {
        return "ellipse (x-y plane), [ra, rb] = " + std::to_string(m_radiusa) + ", " + std::to_string(m_radiusb);
    }
    virtual bool lower(sdfibm_shape_t &o) const override {
        lowerCommon(o, SDFIBM_SHAPE_ELLIPSE);
        o.p[0] = m_radiusa; o.p[1] = m_radiusb; o.p[2] = m_radiusaSQRINV; o.p[3] = m_radiusbSQRINV;
        return true;
    }
    SDFIBM_SHAPE_EVAL_VIA_RECORD()
};

class Ellipsoid : public IShape, public _shapecreator<Ellipsoid> {   // ellipsoid.h:15-54 (never reads `com`)
    scalar m_radiusa, m_radiusb, m_radiusc, m_ia, m_ib, m_ic;

public:
    Ellipsoid(const dictionary &para) {
        m_radiusa = Foam::readScalar(para.lookup("radiusa"));
        m_radiusb = Foam::readScalar(para.lookup("radiusb"));
        m_radiusc = Foam::readScalar(para.lookup("radiusc"));
        m_ia = 1.0 / (m_radiusa * m_radiusa);
        m_ib = 1.0 / (m_radiusb * m_radiusb);
        m_ic = 1.0 / (m_radiusc * m_radiusc);
        m_volume = 4.0 / 3.0 * M_PI * m_radiusa * m_radiusb * m_radiusc;
        m_volumeINV = 1.0 / m_volume;
        detail::diag_moi(*this, 0.2 * m_volume * (m_radiusb * m_radiusb + m_radiusc * m_radiusc),
                         0.2 * m_volume * (m_radiusa * m_radiusa + m_radiusc * m_radiusc),
                         0.2 * m_volume * (m_radiusa * m_radiusa + m_radiusb * m_radiusb));
        m_radiusB = std::max(std::max(m_radiusa, m_radiusb), m_radiusc);
    }
    SHAPETYPENAME("Ellipsoid")
    virtual std::string description() const override {
        return "ellipsoid, [ra, rb, rc] = " + std::to_string(m_radiusa) + ", " + std::to_string(m_radiusb) + ", " + std::to_string(m_radiusc);
    }
    virtual bool lower(sdfibm_shape_t &o) const override {
        lowerCommon(o, SDFIBM_SHAPE_ELLIPSOID);
        o.com[0] = o.com[1] = o.com[2] = 0.0;
        o.p[0] = m_radiusa; o.p[1] = m_radiusb; o.p[2] = m_radiusc; o.p[3] = m_ia; o.p[4] = m_ib; o.p[5] = m_ic;
        return true;
    }
    SDFIBM_SHAPE_EVAL_VIA_RECORD()
};

class Rectangle : public IShape, public _shapecreator<Rectangle> {   // rectangle.h:16-58
    scalar m_radiusa, m_radiusb;

public:
    Rectangle(const dictionary &para) {
        m_radiusa = Foam::readScalar(para.lookup("radiusa"));
        m_radiusb = Foam::readScalar(para.lookup("radiusb"));
        m_com = para.lookupOrDefault("com", vector::zero);
        m_volume = 4.0 * m_radiusa * m_radiusb;
        m_volumeINV = 1.0 / m_volume;
        const scalar tmp = 1.0 / 3.0 * m_volume * (m_radiusa * m_radiusa + m_radiusb * m_radiusb);
        detail::diag_moi(*this, tmp, tmp, tmp);
        m_radiusB = std::max(m_radiusa, m_radiusb);
    }
    SHAPETYPENAME("Rectangle")
    virtual std::string description() const override {
        return "rectangle (x-y plane), [ra, rb] = " + std::to_string(m_radiusa) + ", " + std::to_string(m_radiusb);
    }
    virtual bool lower(sdfibm_shape_t &o) const override {
        lowerCommon(o, SDFIBM_SHAPE_RECTANGLE);
        o.p[0] = m_radiusa; o.p[1] = m_radiusb;
        return true;
    }
    SDFIBM_SHAPE_EVAL_VIA_RECORD()
};

class Box : public IShape, public _shapecreator<Box> {   // box.h:14-57
    scalar m_radiusa, m_radiusb, m_radiusc;

public:
    Box(const dictionary &para) {
        m_radiusa = Foam::readScalar(para.lookup("radiusa"));
        m_radiusb = Foam::readScalar(para.lookup("radiusb"));
        m_radiusc = Foam::readScalar(para.lookup("radiusc"));
        m_com = para.lookupOrDefault("com", vector::zero);
        m_volume = 8.0 * m_radiusa * m_radiusb * m_radiusc;
        m_volumeINV = 1.0 / m_volume;
        detail::diag_moi(*this, 1.0 / 3.0 * m_volume * (m_radiusb * m_radiusb + m_radiusc * m_radiusc),
                         1.0 / 3.0 * m_volume * (m_radiusa * m_radiusa + m_radiusc * m_radiusc),
                         1.0 / 3.0 * m_volume * (m_radiusb * m_radiusb + m_radiusa * m_radiusa));
        m_radiusB = std::max(std::max(m_radiusa, m_radiusb), m_radiusc);
    }
    SHAPETYPENAME("Box")
    virtual std::string description() const override {
        return "Box, [ra, rb, rc] = " + std::to_string(m_radiusa) + ", " + std::to_string(m_radiusb) + ", " + std::to_string(m_radiusc);
    }
    virtual bool lower(sdfibm_shape_t &o) const override {
        lowerCommon(o, SDFIBM_SHAPE_BOX);
        o.p[0] = m_radiusa; o.p[1] = m_radiusb; o.p[2] = m_radiusc;
        return true;
    }
    SDFIBM_SHAPE_EVAL_VIA_RECORD()
};

// circle + one rectangular tail: `thickness` is the tail's HALF width (circle_tail.h:18-61).  A composed shape: the reference's
// U({rectangle(offset(p2d, (ra,0,0)), ra, rb), circle(p2d, r)}) stated as a program (bit for bit the hard-coded tag 7 record,
// tests/test_sdf_programs_cpu.py)
class Circle_Tail : public SdfShape, public _shapecreator<Circle_Tail> {
protected:
    scalar m_radius, m_ratio, m_radiusb, m_radiusSQR, m_radiusa;

public:
    Circle_Tail(const dictionary &para) {
        m_radius = Foam::readScalar(para.lookup("radius"));
        m_ratio = Foam::readScalar(para.lookup("ratio"));
        m_radiusb = Foam::readScalar(para.lookup("thickness"));
        m_com = para.lookupOrDefault("com", vector::zero);
        m_radiusSQR = m_radius * m_radius;
        m_radiusa = (m_ratio + 1) * 0.5 * m_radius;
        m_volume = M_PI * m_radiusSQR;   // the tail is ignored in the mass properties
        m_volumeINV = 1.0 / m_volume;
        const scalar tmp = 0.5 * m_volume * m_radiusSQR;
        detail::diag_moi(*this, tmp, tmp, tmp);
        m_radiusB = 2 * m_radiusa;
        m_program.point2d().circle(m_radius)                                                        // circle_tail.h:47,55
            .point2d().offset(vector(m_radiusa, 0.0, 0.0)).rectangle(m_radiusa, m_radiusb).unite();   // :46,56-57 and U (:50,59)
        // tail box [0, 2 ra] x [-rb, rb]
        setBounds(std::max(m_radius, std::sqrt(4 * m_radiusa * m_radiusa + m_radiusb * m_radiusb)) * 1.001, m_radius, true);
    }
    SHAPETYPENAME("Circle_Tail")
    virtual std::string description() const override { return "Circle_Tail (x-y plane), r = " + std::to_string(m_radius); }
    virtual bool lower(sdfibm_shape_t &o) const override {   // the hard-coded record (C-ABI tag 7), kept for hosts that fill records themselves
        lowerCommon(o, SDFIBM_SHAPE_CIRCLE_TAIL);
        o.p[0] = m_radius; o.p[1] = m_radiusSQR; o.p[2] = m_radiusa; o.p[3] = m_radiusb;
        return true;
    }
};

// circle + two tails at +-30 degrees: the dictionary `thickness` is halved here (circle_twotail.h:18-66); composed like Circle_Tail
class Circle_TwoTail : public SdfShape, public _shapecreator<Circle_TwoTail> {
    scalar m_radius, m_ratio, m_radiusb, m_radiusSQR, m_radiusa;

public:
    Circle_TwoTail(const dictionary &para) {
        m_radius = Foam::readScalar(para.lookup("radius"));
        m_ratio = Foam::readScalar(para.lookup("ratio"));
        m_radiusb = Foam::readScalar(para.lookup("thickness")) * 0.5;
        m_com = para.lookupOrDefault("com", vector::zero);
        m_radiusSQR = m_radius * m_radius;
        m_radiusa = (m_ratio + 1) * 0.5 * m_radius;
        m_volume = M_PI * m_radiusSQR;
        m_volumeINV = 1.0 / m_volume;
        const scalar tmp = 0.5 * m_volume * m_radiusSQR;
        detail::diag_moi(*this, tmp, tmp, tmp);
        m_radiusB = 2 * m_radiusa;
        const vector shift(m_radiusa, 0, 0);
        m_program.point2d().circle(m_radius)                                                 // dc  (circle_twotail.h:47,57)
            .point2d().rot30().offset(shift).rectangle(m_radiusa, m_radiusb).unite()             // d1  (:48-49,58-59)
            .point2d().flipy().rot30().offset(shift).rectangle(m_radiusa, m_radiusb).unite();    // d2  (:50-51,60-61); U({dc,d1,d2}) (:53,63)
        setBounds(std::max(m_radius, std::sqrt(4 * m_radiusa * m_radiusa + m_radiusb * m_radiusb) * 1.001), m_radius, true);   // rot30's literal is not unitary
    }
    SHAPETYPENAME("Circle_TwoTail")
    virtual std::string description() const override { return "Circle_TwoTail (x-y plane), r = " + std::to_string(m_radius); }
    virtual bool lower(sdfibm_shape_t &o) const override {
        lowerCommon(o, SDFIBM_SHAPE_CIRCLE_TWOTAIL);
        o.p[0] = m_radius; o.p[1] = m_radiusSQR; o.p[2] = m_radiusa; o.p[3] = m_radiusb;
        return true;
    }
};

// half space local y < 0; not finite: zero volume, so mass_inv = 0 and moi_inv = I / rho (plane.h:13-28, solid.h:95-103)
class Plane : public IShape, public _shapecreator<Plane> {
public:
    Plane(const dictionary &) { finite = false; }
    SHAPETYPENAME("Plane")
    virtual std::string description() const override { return "Plane (x-z plane)"; }
    virtual bool lower(sdfibm_shape_t &o) const override {
        lowerCommon(o, SDFIBM_SHAPE_PLANE);
        return true;
    }
    SDFIBM_SHAPE_EVAL_VIA_RECORD()
};

} // namespace sdfibm
