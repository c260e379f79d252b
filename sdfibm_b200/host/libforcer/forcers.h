// forcers.h — external force/torque generators (reference src/libforcer/{iforcer,forcerfactory,constant,spring,
// magnetic}.h): generate(t, x, v, q, omega) -> (F, T), blended 1.5 new - 0.5 old by Solid::applyForcer.
#pragma once
#include <cmath>
#include <memory>
#include <utility>

#include "../genericfactory.h"
#include "../types.h"

namespace sdfibm {
namespace forcer {

#define FORCERTYPENAME(name)                       \
    static std::string typeName() { return name; } \
    static bool added;

class IForcer;
template <typename T>
class _creator {
public:
    static std::unique_ptr<IForcer> create(const dictionary &para) { return std::make_unique<T>(para); }
};

class IForcer {
public:
    using Force = std::pair<vector, vector>;
    IForcer() = default;
    virtual ~IForcer() = default;
    virtual Force generate(const scalar &time, const vector &position, const vector &velocity, const quaternion &orientation,
                           const vector &omega) = 0;
    virtual std::string description() const = 0;
};

MAKESPECIALFACTORY(Forcer, IForcer, dictionary);
#define REGISTERFORCE(m) bool sdfibm::m::added = sdfibm::forcer::ForcerFactory::add(sdfibm::m::typeName(), sdfibm::m::create);

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
