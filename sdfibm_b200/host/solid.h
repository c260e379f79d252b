// solid.h — the rigid-body side of the cloud as a struct of arrays.
//
// The reference keeps one `Solid` object per body (src/solid.h) and walks the objects five times per sub-iteration
// (src/solidcloud.cpp:521-562).  Here the cloud owns ONE SolidStates: an array per quantity, and the integrator is five batched
// passes over those arrays — clearLoads / applyForcers / blendFluidLoads / addBuoyantWeight / advance — the layout the device
// records are packed from (pack()) and the per-solid sums are unpacked into (setFluidLoads()).  The arithmetic of every pass is
// the reference's, statement for statement where rounding could differ (the AB2 blend 1.5 new - 0.5 old, the world-frame inverse
// inertia R I^-1 R^T, the mid-point position / quaternion update, solid.h:148-202), so trajectories are bit-identical to the
// reference's compiled Solid (tests/test_oracle_vs_reference.py, tests/test_oracle_fuzz_vs_reference.py).
// `Solid` is a read-only view of one row (what SolidCloud::operator[] hands out).
#pragma once
#include <ostream>
#include <vector>

#include "libforcer/iforcer.h"
#include "libmaterial/imaterial.h"
#include "libmotion/imotion.h"
#include "libshape/ishape.h"
#include "types.h"

namespace sdfibm {

class SolidStates {
public:
    // ---- kinematic state: what the coupling kernels read ----
    std::vector<vector> x;          // centre
    std::vector<quaternion> q;      // orientation
    std::vector<vector> v, w;       // velocity, angular velocity
    // ---- loads of the current sub-iteration and the samples the two-level blend needs ----
    std::vector<vector> F, T;
    std::vector<vector> F_fluid, T_fluid, F_fluid_prev, T_fluid_prev;
    std::vector<vector> F_ext, T_ext, F_ext_prev, T_ext_prev;
    std::vector<char> fluid_primed, ext_primed;   // a previous sample exists (the first blend uses the new one twice)
    // ---- inertia = shape x material (solid.h:95-103) ----
    std::vector<scalar> mass, mass_inv;
    std::vector<tensor> inertia_inv_body;
    // ---- plugins (owned by the cloud's libraries) ----
    std::vector<IShape *> shape;
    std::vector<IMotion *> motion;
    std::vector<forcer::IForcer *> forcer;
    std::vector<IMaterial *> material;

    size_t size() const { return x.size(); }
    bool empty() const { return x.empty(); }
    void reserve(size_t n) { x.reserve(n); q.reserve(n); v.reserve(n); w.reserve(n); }

    // a new body at rest; orientation from XYZ Euler angles in radians (solidcloud.cpp:117: Foam::quaternion(XYZ, euler))
    size_t add(const vector &centre, const vector &euler_xyz_rad) {
        const vector zero = vector::zero;
        x.push_back(centre);
        q.push_back(quaternion(quaternion::XYZ, euler_xyz_rad));
        v.push_back(zero); w.push_back(zero);
        for (auto *a : {&F, &T, &F_fluid, &T_fluid, &F_fluid_prev, &T_fluid_prev, &F_ext, &T_ext, &F_ext_prev, &T_ext_prev}) a->push_back(zero);
        fluid_primed.push_back(0); ext_primed.push_back(0);
        mass.push_back(0); mass_inv.push_back(0);
        inertia_inv_body.push_back(tensor::I);
        shape.push_back(nullptr); motion.push_back(nullptr); forcer.push_back(nullptr); material.push_back(nullptr);
        return x.size() - 1;
    }
    void setShapeAndMaterial(size_t i, IShape *sh, IMaterial *mat) {
        shape[i] = sh;
        material[i] = mat;
        const scalar rho = mat->getRho();
        mass[i] = sh->m_volume * rho;
        mass_inv[i] = sh->m_volumeINV / rho;
        inertia_inv_body[i] = sh->m_moiINV / rho;
    }

    // ---- the batched passes of one sub-iteration (src/solidcloud.cpp:530-556) ----
    void clearLoads() {
        for (size_t i = 0; i < size(); ++i) { F[i] = vector::zero; T[i] = vector::zero; }
    }
    void applyForcers(scalar time) {   // solid.h:148-163
        for (size_t i = 0; i < size(); ++i) {
            if (!forcer[i]) continue;
            const auto ft = forcer[i]->generate(time, x[i], v[i], q[i], w[i]);
            F_ext[i] = ft.first;
            T_ext[i] = ft.second;
            if (!ext_primed[i]) { F_ext_prev[i] = F_ext[i]; T_ext_prev[i] = T_ext[i]; ext_primed[i] = 1; }
            F[i] += (1.5 * F_ext[i] - 0.5 * F_ext_prev[i]);
            T[i] += (1.5 * T_ext[i] - 0.5 * T_ext_prev[i]);
        }
    }
    void blendFluidLoads() {   // solid.h:164-174
        for (size_t i = 0; i < size(); ++i) {
            if (!fluid_primed[i]) { F_fluid_prev[i] = F_fluid[i]; T_fluid_prev[i] = T_fluid[i]; fluid_primed[i] = 1; }
            F[i] += (1.5 * F_fluid[i] - 0.5 * F_fluid_prev[i]);
            T[i] += (1.5 * T_fluid[i] - 0.5 * T_fluid_prev[i]);
        }
    }
    void addBuoyantWeight(const vector &gravity, scalar rho_fluid) {   // solidcloud.cpp:466-475
        for (size_t i = 0; i < size(); ++i) {
            const scalar rhos = material[i]->getRho();
            const vector gprime = ((rhos - rho_fluid) / rhos) * gravity;
            F[i] += mass[i] * gprime;
        }
    }
    void addLoads(const double *ft6) {   // per-solid (F, T) rows, e.g. the contact forces of the collision step
        for (size_t i = 0; i < size(); ++i) {
            F[i] += vector(ft6[6 * i], ft6[6 * i + 1], ft6[6 * i + 2]);
            T[i] += vector(ft6[6 * i + 3], ft6[6 * i + 4], ft6[6 * i + 5]);
        }
    }
    void advance(scalar time, scalar dt) {   // Solid::move, solid.h:181-202
        for (size_t i = 0; i < size(); ++i) {
            const vector v_old = v[i], w_old = w[i];
            v[i] += F[i] * mass_inv[i] * dt;
            const tensor R = q[i].R();
            const tensor inertia_inv_world = R & inertia_inv_body[i] & R.T();
            w[i] += (inertia_inv_world & T[i]) * dt;
            if (motion[i] != nullptr) motion[i]->constraint(time, v[i], w[i]);
            x[i] += 0.5 * (v[i] + v_old) * dt;
            q[i] += 0.5 * quaternion(0.5 * (w[i] + w_old)) * q[i] * dt;
            q[i].normalise();
        }
    }
    void rememberLoads() {   // storeOldForce, solid.h:140-147
        F_fluid_prev = F_fluid; T_fluid_prev = T_fluid;
        F_ext_prev = F_ext; T_ext_prev = T_ext;
    }
    void setFluidLoads(const double *ft6) {   // solidcloud.cpp:432
        for (size_t i = 0; i < size(); ++i) {
            F_fluid[i] = vector(ft6[6 * i], ft6[6 * i + 1], ft6[6 * i + 2]);
            T_fluid[i] = vector(ft6[6 * i + 3], ft6[6 * i + 4], ft6[6 * i + 5]);
        }
    }

    // ---- the device records (include/sdfibm_b200.h sdfibm_solid_t) ----
    void packOne(size_t i, sdfibm_solid_t &r, int shape_index) const {
        r.pos[0] = x[i].x(); r.pos[1] = x[i].y(); r.pos[2] = x[i].z();
        r.quat[0] = q[i].w(); r.quat[1] = q[i].v().x(); r.quat[2] = q[i].v().y(); r.quat[3] = q[i].v().z();
        r.vel[0] = v[i].x(); r.vel[1] = v[i].y(); r.vel[2] = v[i].z();
        r.omega[0] = w[i].x(); r.omega[1] = w[i].y(); r.omega[2] = w[i].z();
        r.shape = shape_index;
        r.pad_ = 0;
    }
    // shape_index: one index per solid, or (n == 0) `every` for all of them
    void pack(std::vector<sdfibm_solid_t> &out, const std::vector<int> &shape_index, int every = -1) const {
        out.resize(size());
        for (size_t i = 0; i < size(); ++i) packOne(i, out[i], shape_index.empty() ? every : shape_index[i]);
    }

    // ---- cloud.out rows: 18 columns in 3-D, 9 in 2-D (src/solid.cpp:5-29) ----
    void writeRow(std::ostream &os, size_t i, bool two_d) const {
        const vector e = q[i].eulerAngles(quaternion::XYZ);
        auto three = [&os](const vector &a, const char *end) { os << a.x() << ' ' << a.y() << ' ' << a.z() << end; };
        auto two = [&os](const vector &a) { os << a.x() << ' ' << a.y() << ' '; };
        if (!two_d) {
            three(x[i], " "); three(v[i], " "); three(F[i], " "); three(e, " "); three(w[i], " "); three(T[i], "");
        } else {
            two(x[i]); two(v[i]); two(F[i]);
            os << e.z() << ' ' << w[i].z() << ' ' << T[i].z();
        }
    }
};

// one row of the cloud, read-only (SolidCloud::operator[])
class Solid {
    const SolidStates *s_;
    size_t i_;

public:
    Solid(const SolidStates &s, size_t i) : s_(&s), i_(i) {}
    label getID() const { return (label)i_; }
    const vector &getCenter() const { return s_->x[i_]; }
    const vector &getVelocity() const { return s_->v[i_]; }
    const vector &getOmega() const { return s_->w[i_]; }
    const vector &getForce() const { return s_->F[i_]; }
    const vector &getTorque() const { return s_->T[i_]; }
    const quaternion &getOrientation() const { return s_->q[i_]; }
    const vector &getFluidForce() const { return s_->F_fluid[i_]; }
    const vector &getFluidTorque() const { return s_->T_fluid[i_]; }
    IMotion *getMotion() const { return s_->motion[i_]; }
    IShape *getShape() const { return s_->shape[i_]; }
    IMaterial *getMaterial() const { return s_->material[i_]; }
    scalar getRadiusB() const { return s_->shape[i_]->getRadiusB(); }
    bool isFinite() const { return s_->shape[i_]->finite; }
    scalar getMass() const { return s_->mass[i_]; }
    bool phi01(const vector &p) const { return s_->shape[i_]->phi01(p, {s_->x[i_], s_->q[i_]}); }
    scalar phi(const vector &p) const { return s_->shape[i_]->phi(p, {s_->x[i_], s_->q[i_]}); }
    vector evalPointVelocity(const vector &p) const { return s_->v[i_] + (s_->w[i_] ^ (p - s_->x[i_])); }
    void toRecord(sdfibm_solid_t &r, int shape_index) const { s_->packOne(i_, r, shape_index); }
};

} // namespace sdfibm
