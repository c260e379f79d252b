// solid.h — rigid-body state and integrator of one solid (reference src/solid.h:12-217, src/solid.cpp:5-29).
// The per-step coupling reads (center, orientation, velocity, omega, shape) through toRecord(); everything else —
// force blending, constraints, the move() integrator — is O(N_solids) host work and stays here.
#pragma once
#include <ostream>

#include "libforcer/iforcer.h"
#include "libmaterial/imaterial.h"
#include "libmotion/imotion.h"
#include "libshape/ishape.h"
#include "types.h"

namespace sdfibm {

class Solid {
protected:
    label id;
    label hostid{-1};
    vector center;
    quaternion orientation;
    vector velocity{vector::zero};
    vector omega{vector::zero};
    vector force{vector::zero};
    vector torque{vector::zero};
    vector fluid_force{vector::zero};
    vector fluid_torque{vector::zero};
    vector fluid_force_old{vector::zero};
    vector fluid_torque_old{vector::zero};
    bool first_fluid_step{true};
    vector forcer_force{vector::zero};
    vector forcer_torque{vector::zero};
    vector forcer_force_old{vector::zero};
    vector forcer_torque_old{vector::zero};
    bool first_forcer_step{true};
    IMotion *ptr_motion{nullptr};
    IShape *ptr_shape{nullptr};
    forcer::IForcer *ptr_forcer{nullptr};
    IMaterial *ptr_material{nullptr};
    scalar mass{0};
    scalar mass_inv{0};
    tensor moi_inv{tensor::I};

public:
    Solid(label solid_id, const vector &solid_center, const quaternion &solid_quaternion)
        : id(solid_id), center(solid_center), orientation(solid_quaternion) {}

    label getID() const { return id; }
    const vector &getCenter() const { return center; }
    const vector &getVelocity() const { return velocity; }
    const vector &getOmega() const { return omega; }
    const vector &getForce() const { return force; }
    const vector &getTorque() const { return torque; }
    const quaternion &getOrientation() const { return orientation; }
    const vector &getFluidForce() const { return fluid_force; }
    const vector &getFluidTorque() const { return fluid_torque; }

    void setCenter(const vector &c) { center = c; }
    void setVelocity(const vector &v) { velocity = v; }
    void setOmega(const vector &o) { omega = o; }
    void setForce(const vector &f) { force = f; }
    void setTorque(const vector &t) { torque = t; }
    void setOrientation(const vector &angles) { orientation = quaternion(quaternion::XYZ, angles); }   // radians, XYZ

    IMotion *getMotion() const { return ptr_motion; }
    IShape *getShape() const { return ptr_shape; }
    IMaterial *getMaterial() const { return ptr_material; }
    scalar getRadiusB() const { return ptr_shape->getRadiusB(); }
    bool isFinite() const { return ptr_shape->finite; }
    scalar getMass() const { return mass; }

    void setShape(IShape *shape) { ptr_shape = shape; }
    void setForcer(forcer::IForcer *f) { ptr_forcer = f; }
    void setMaterial(IMaterial *material) {   // object properties = shape x material (solid.h:95-103)
        ptr_material = material;
        const scalar rho = ptr_material->getRho();
        mass = ptr_shape->m_volume * rho;
        mass_inv = ptr_shape->m_volumeINV / rho;
        moi_inv = ptr_shape->m_moiINV / rho;
    }
    void setMotion(IMotion *m) { ptr_motion = m; }
    void unsetMotion() { ptr_motion = nullptr; }

    bool phi01(const vector &p) const { return ptr_shape->phi01(p, {center, orientation}); }
    scalar phi(const vector &p) const { return ptr_shape->phi(p, {center, orientation}); }
    vector evalPointVelocity(const vector &p) const { return velocity + (omega ^ (p - center)); }

    void addAcceleration(const vector &acc) { force += mass * acc; }
    void clearForceAndTorque() {
        force = vector::zero;
        torque = vector::zero;
    }
    void setFluidForceAndTorque(const vector &f, const vector &t) {
        fluid_force = f;
        fluid_torque = t;
    }
    void storeOldForce() {
        fluid_force_old = fluid_force;
        fluid_torque_old = fluid_torque;
        forcer_force_old = forcer_force;
        forcer_torque_old = forcer_torque;
    }
    void applyForcer(scalar &time) {   // solid.h:148-163
        if (!ptr_forcer) return;
        auto ft = ptr_forcer->generate(time, center, velocity, orientation, omega);
        forcer_force = ft.first;
        forcer_torque = ft.second;
        if (first_forcer_step) {
            forcer_force_old = forcer_force;
            forcer_torque_old = forcer_torque;
            first_forcer_step = false;
        }
        force += (1.5 * forcer_force - 0.5 * forcer_force_old);
        torque += (1.5 * forcer_torque - 0.5 * forcer_torque_old);
    }
    void addMidFluidForceAndTorque() {   // solid.h:164-174
        if (first_fluid_step) {
            fluid_force_old = fluid_force;
            fluid_torque_old = fluid_torque;
            first_fluid_step = false;
        }
        force += (1.5 * fluid_force - 0.5 * fluid_force_old);
        torque += (1.5 * fluid_torque - 0.5 * fluid_torque_old);
    }
    void addForceAndTorque(const vector &f, const vector &t) {
        force += f;
        torque += t;
    }
    void move(const scalar &time, const scalar &dt) {   // solid.h:181-202
        const vector velocity_old = velocity;
        const vector omega_old = omega;
        velocity += force * mass_inv * dt;
        const tensor R = orientation.R();
        const tensor moi_inv_world = R & moi_inv & R.T();
        omega += (moi_inv_world & torque) * dt;
        if (ptr_motion != nullptr) ptr_motion->constraint(time, velocity, omega);
        center += 0.5 * (velocity + velocity_old) * dt;
        orientation += 0.5 * quaternion(0.5 * (omega + omega_old)) * orientation * dt;
        orientation.normalise();
    }

    // the rigid-body record the device path reads (include/sdfibm_b200.h sdfibm_solid_t)
    void toRecord(sdfibm_solid_t &r, int shape_index) const {
        r.pos[0] = center.x(); r.pos[1] = center.y(); r.pos[2] = center.z();
        r.quat[0] = orientation.w(); r.quat[1] = orientation.v().x(); r.quat[2] = orientation.v().y(); r.quat[3] = orientation.v().z();
        r.vel[0] = velocity.x(); r.vel[1] = velocity.y(); r.vel[2] = velocity.z();
        r.omega[0] = omega.x(); r.omega[1] = omega.y(); r.omega[2] = omega.z();
        r.shape = shape_index;
        r.pad_ = 0;
    }

    friend std::ostream &operator<<(std::ostream &os, const Solid &s) {   // 18 columns, 3-D (solid.cpp:5-17)
        vector v;
        v = s.getCenter();   os << v.x() << ' ' << v.y() << ' ' << v.z() << ' ';
        v = s.getVelocity(); os << v.x() << ' ' << v.y() << ' ' << v.z() << ' ';
        v = s.getForce();    os << v.x() << ' ' << v.y() << ' ' << v.z() << ' ';
        v = s.getOrientation().eulerAngles(quaternion::XYZ);
        os << v.x() << ' ' << v.y() << ' ' << v.z() << ' ';
        v = s.getOmega();    os << v.x() << ' ' << v.y() << ' ' << v.z() << ' ';
        v = s.getTorque();   os << v.x() << ' ' << v.y() << ' ' << v.z();
        return os;
    }
    friend void write2D(std::ostream &os, const Solid &s) {   // 9 columns, 2-D (solid.cpp:18-29)
        vector v;
        v = s.getCenter();   os << v.x() << ' ' << v.y() << ' ';
        v = s.getVelocity(); os << v.x() << ' ' << v.y() << ' ';
        v = s.getForce();    os << v.x() << ' ' << v.y() << ' ';
        v = s.getOrientation().eulerAngles(quaternion::XYZ);
        os << v.z() << ' ';
        v = s.getOmega();    os << v.z() << ' ';
        v = s.getTorque();   os << v.z();
    }
};

} // namespace sdfibm
