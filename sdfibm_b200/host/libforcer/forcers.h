// forcers.h — the three built-in forcers (reference src/libforcer/{constant,spring,magnetic}.h).
#pragma once
#include <cmath>

#include "forcerfactory.h"

namespace sdfibm {
namespace forcer {

class Constant : public IForcer, public _creator<Constant> {   // constant.h:14-33
    vector force, torque;

public:
    FORCERTYPENAME("Constant")
    Constant(const dictionary &para) {
        force = para.lookup("force");
        torque = para.lookup("torque");
    }
    virtual Force generate(const scalar &, const vector &, const vector &, const quaternion &, const vector &) override final {
        return {force, torque};
    }
    virtual std::string description() const override { return "forcer with const force and torque"; }
};

class Spring : public IForcer, public _creator<Spring> {   // spring.h:12-46
    vector pivot;
    scalar k, l;

public:
    FORCERTYPENAME("Spring")
    Spring(const dictionary &para) {
        pivot = para.lookup("pivot");
        k = Foam::readScalar(para.lookup("k"));
        l = Foam::readScalar(para.lookup("l"));
    }
    virtual Force generate(const scalar &, const vector &position, const vector &, const quaternion &, const vector &) override final {
        const vector r = position - pivot;
        vector force = vector::zero;
        if (Foam::mag(r) > SMALL) force = -k * r * (1.0 - l / Foam::mag(r));
        return {force, vector::zero};
    }
    virtual std::string description() const override { return "Spring forcer with a pivot, stiffness (k), and rest length (l)"; }
};

class Magnetic : public IForcer, public _creator<Magnetic> {   // magnetic.h:12-47
    vector direction;
    scalar A, w;

public:
    FORCERTYPENAME("Magnetic")
    Magnetic(const dictionary &para) {
        direction = para.lookup("direction");
        A = Foam::readScalar(para.lookup("A"));
        w = Foam::readScalar(para.lookup("w"));
    }
    virtual Force generate(const scalar &time, const vector &, const vector &, const quaternion &orientation, const vector &) override final {
        const vector m_original = (1.0) * vector(0.0, 0.0, 1.0);   // unit magnetic moment along body z
        const vector B = A * std::cos(w * time) * direction;
        const vector m = Foam::conjugate(orientation).transform(m_original);
        return {vector::zero, m ^ B};
    }
    virtual std::string description() const override { return "Magnetic forcer: A*cos(omega*t)*direction"; }
};

} // namespace forcer
} // namespace sdfibm

#ifdef SDFIBM_REGISTER_BUILTINS
REGISTERFORCE(forcer::Constant)
REGISTERFORCE(forcer::Spring)
REGISTERFORCE(forcer::Magnetic)
#endif
