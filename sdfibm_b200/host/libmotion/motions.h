// motions.h — the seven motion plugins (reference src/libmotion/motion*.h).
#pragma once
#include <cmath>
#include <stdexcept>

#include "imotion.h"

namespace sdfibm {

// `mask b` + six 0/1 flags for (vx vy vz wx wy wz): component-wise multiply (motion01mask.h:21-64)
class Motion01Mask : public IMotion, public _creator<Motion01Mask> {
    vector vmask, omask;

public:
    TYPENAME("Motion01Mask")
    Motion01Mask(const dictionary &para) {
        const std::string mask = Foam::word(para.lookup("mask"));
        if (mask.size() != 7 || mask[0] != 'b') throw std::runtime_error("Motion01Mask: mask must be b followed by six 0/1 digits");
        for (int i = 1; i <= 6; ++i)
            if (mask[i] != '0' && mask[i] != '1') throw std::runtime_error("Motion01Mask: mask element is not 0 or 1");
        vmask = vector(mask[1] == '0' ? 0 : 1, mask[2] == '0' ? 0 : 1, mask[3] == '0' ? 0 : 1);
        omask = vector(mask[4] == '0' ? 0 : 1, mask[5] == '0' ? 0 : 1, mask[6] == '0' ? 0 : 1);
    }
    virtual void constraint(const scalar &, vector &velocity, vector &omega) override final {
        velocity = Foam::cmptMultiply(velocity, vmask);
        omega = Foam::cmptMultiply(omega, omask);
    }
    virtual std::string description() const override { return "general (0|1){6} motion mask"; }
};

// fixed position, constant spin about z with the given period (motion000002.h:17-40)
class Motion000002 : public IMotion, public _creator<Motion000002> {
    scalar m_period, m_omega;

public:
    TYPENAME("Motion000002")
    Motion000002(const dictionary &para) {
        m_period = Foam::readScalar(para.lookup("period"));
        m_omega = 2 * M_PI / m_period;
    }
    virtual void constraint(const scalar &, vector &velocity, vector &omega) override final {
        velocity = vector::zero;
        omega = vector::zero;
        omega[2] = m_omega;
    }
    virtual std::string description() const override { return "fixed centre, constant rotation about z"; }
};

// free in x, y; vz = 0; constant spin about z (motion110002.h:17-40)
class Motion110002 : public IMotion, public _creator<Motion110002> {
    scalar m_period, m_omega;

public:
    TYPENAME("Motion110002")
    Motion110002(const dictionary &para) {
        m_period = Foam::readScalar(para.lookup("period"));
        m_omega = 2 * M_PI / m_period;
    }
    virtual void constraint(const scalar &, vector &velocity, vector &omega) override final {
        velocity[2] = 0.0;
        omega = vector::zero;
        omega[2] = m_omega;
    }
    virtual std::string description() const override { return "free in-plane translation, constant rotation about z"; }
};

// prescribed constant velocity (u v w), no rotation (motion222000.h:17-39)
class Motion222000 : public IMotion, public _creator<Motion222000> {
    scalar m_u, m_v, m_w;

public:
    TYPENAME("Motion222000")
    Motion222000(const dictionary &para) {
        m_u = Foam::readScalar(para.lookup("u"));
        m_v = Foam::readScalar(para.lookup("v"));
        m_w = Foam::readScalar(para.lookup("w"));
    }
    virtual void constraint(const scalar &, vector &velocity, vector &omega) override final {
        velocity = vector(m_u, m_v, m_w);
        omega = vector::zero;
    }
    virtual std::string description() const override { return "constant translation, no rotation"; }
};

// v = A w cos(w t) dir, no rotation (motionsinedirectional.h:17-48)
class MotionSineDirectional : public IMotion, public _creator<MotionSineDirectional> {
    scalar m_amplitude, m_period, m_omega;
    vector m_direction;

public:
    TYPENAME("MotionSineDirectional")
    MotionSineDirectional(const dictionary &para) {
        m_amplitude = Foam::readScalar(para.lookup("amplitude"));
        m_period = Foam::readScalar(para.lookup("period"));
        m_omega = 2 * M_PI / m_period;
        m_direction = para.lookup("direction");
        if (Foam::magSqr(m_direction) > 1.001) throw std::runtime_error("MotionSineDirectional: direction vector not normalized!");
    }
    virtual void constraint(const scalar &time, vector &velocity, vector &omega) override final {
        velocity = m_amplitude * m_omega * std::cos(m_omega * time) * m_direction;
        omega = vector::zero;
    }
    virtual std::string description() const override { return "linear oscillation along a direction"; }
};

// orbit about the origin with the given radius/period/phase, self spin `selfom` about z (motionrotor.h:17-47)
class MotionRotor : public IMotion, public _creator<MotionRotor> {
    scalar m_period, m_radius, m_theta0, m_selfom, m_omega;

public:
    TYPENAME("MotionRotor")
    MotionRotor(const dictionary &para) {
        m_period = Foam::readScalar(para.lookup("period"));
        m_radius = Foam::readScalar(para.lookup("radius"));
        m_theta0 = Foam::readScalar(para.lookup("theta0"));
        m_selfom = Foam::readScalar(para.lookup("selfom"));
        m_omega = 2 * M_PI / m_period;
    }
    virtual void constraint(const scalar &time, vector &velocity, vector &omega) override final {
        velocity = vector::zero;
        velocity[0] = -m_radius * m_omega * std::sin(m_omega * time + m_theta0);
        velocity[1] = m_radius * m_omega * std::cos(m_omega * time + m_theta0);
        omega = vector::zero;
        omega[2] = m_selfom;
    }
    virtual std::string description() const override { return "orbital motion about the origin with self rotation"; }
};

// gate: vy = -1 for t mod 5 in (1,2), +1 for t mod 5 in (3,4) (motionopenclose.h:24-40)
class MotionOpenClose : public IMotion, public _creator<MotionOpenClose> {
public:
    TYPENAME("MotionOpenClose")
    MotionOpenClose(const dictionary &) {}
    virtual void constraint(const scalar &time, vector &velocity, vector &omega) override final {
        const scalar t = std::fmod(time, 5.0);
        scalar v = 0.0;
        if (t > 1 && t < 2) v = -1.0;
        if (t > 3 && t < 4) v = 1.0;
        velocity = vector(0, v, 0);
        omega = vector::zero;
    }
    virtual std::string description() const override { return "model the open-close operation of a gate"; }
};

} // namespace sdfibm

#ifdef SDFIBM_REGISTER_BUILTINS
REGISTERMOTION(Motion000002)
REGISTERMOTION(Motion110002)
REGISTERMOTION(Motion222000)
REGISTERMOTION(MotionSineDirectional)
REGISTERMOTION(Motion01Mask)
REGISTERMOTION(MotionRotor)
REGISTERMOTION(MotionOpenClose)
#endif
