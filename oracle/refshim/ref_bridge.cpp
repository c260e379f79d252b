// ref_bridge.cpp — TEST INFRASTRUCTURE.  A C entry over the reference's OWN CellEnumerator / GeometricTools / shape classes
// (compiled unmodified from /root/reference/src through oracle/refshim): for each solid, the three candidate lists and the
// clipped volume fraction field, following solidFluidInteract (reference src/solidcloud.cpp:361-410) up to alpha.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <unistd.h>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>

#include "libshape/shapefactory.h"
#include "cellenumerator.h"
#include "geometrictools.h"
#include "solid.h"
#include "libcollision/ugrid.h"
#include "libcollision/collision.h"
#include "libmotion/motionfactory.h"
#include "libforcer/forcerfactory.h"

using namespace sdfibm;

extern "C" {
struct ref_mesh;
}
static void fill_mesh(Foam::fvMesh &mesh, const ref_mesh *m);

extern "C" {

struct ref_mesh {
    int32_t n_cells, n_points, n_faces;
    const double *points, *cc, *V, *Cf, *Sf;
    const int32_t *cp_off, *cp, *cf_off, *cf, *fp_off, *fp, *nb_off, *nb;
};

// dict_text[s]: OpenFOAM-style entries of the shape of solid s ("type Sphere; radius 5;"), pos[3s], quat[4s] (w, x, y, z), seed[s]
// = nearest cell to the centre (meshSearch::findNearestCell is OpenFOAM's, not the reference's).  Outputs: list_off[3n+1],
// list_cells (capacity cap), As[n_cells] = min(sum alpha, 1).  Returns the number of list entries, or -1.
int64_t ref_interact(const ref_mesh *m, int n_solids, const char *const *dict_text, const double *pos, const double *quat,
                     const int32_t *seed, int two_d, int32_t *list_off, int32_t *list_cells, int64_t cap, double *As) {
    try {
        Foam::fvMesh mesh;
        fill_mesh(mesh, m);
        GeometricTools geo(mesh);
        for (int c = 0; c < m->n_cells; ++c) As[c] = 0.0;
        int64_t n_out = 0;
        list_off[0] = 0;
        for (int s = 0; s < n_solids; ++s) {
            char tmpl[] = "/tmp/sdfibm_ref_dict_XXXXXX";
            const int fd = mkstemp(tmpl);
            if (fd < 0) return -3;
            close(fd);
            { std::ofstream os(tmpl); os << dict_text[s] << "\n"; }
            Foam::dictionary d = Foam::dictionary::fromFile(tmpl);
            std::remove(tmpl);
            const std::string type = std::string(d.lookup("type"));
            std::unique_ptr<IShape> shape = ShapeFactory::create(type, d);
            Solid solid(s, Foam::vector(pos[3 * s], pos[3 * s + 1], pos[3 * s + 2]),
                        Foam::quaternion(quat[4 * s], Foam::vector(quat[4 * s + 1], quat[4 * s + 2], quat[4 * s + 3])));
            solid.setShape(shape.get());
            CellEnumerator ce(mesh, [&](const vector &p) { return solid.phi01(p); }, seed[s]);
            const CellEnumerator::IntersectionSet &is_ = ce.intersect();
            geo.clearCache();
            const CellEnumerator::CELL_TYPE types[3] = {CellEnumerator::ALL_INSIDE, CellEnumerator::CENTER_INSIDE, CellEnumerator::CENTER_OUTSIDE};
            for (int t = 0; t < 3; ++t) {
                auto it = is_.find(types[t]);
                if (it != is_.end())
                    for (size_t icell : it->second) {
                        if (n_out >= cap) return -1;
                        list_cells[n_out++] = (int32_t)icell;
                        As[icell] += (t == 0) ? 1.0 : geo.calcCellVolume((label)icell, solid, two_d != 0) / mesh.cv[icell];   // solidcloud.cpp:376-410
                    }
                list_off[3 * s + t + 1] = (int32_t)n_out;
            }
        }
        for (int c = 0; c < m->n_cells; ++c) As[c] = std::min(As[c], 1.0);   // checkAlpha, :564-570
        return n_out;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_interact: %s\n", e.what());
        return -2;
    }
}

// The collision step: UGrid broad phase and the narrow-phase table are the reference's (src/libcollision/*, compiled unmodified);
// the loop and the force law around them are the caller's lines, restated from src/solidcloud.cpp:477-519.
// pairs[2*cap] receives the (first, second) pairs in generateCollisionPairs order; ft[6n] is accumulated into.
int64_t ref_collide(const double *bmin, const double *bmax, double delta, int n_solids, const char *const *dict_text, const double *pos,
                    const double *quat, int32_t *pairs, int64_t cap, double *ft) {
    try {
        static bool table = false;
        if (!table) { InitCollisionFuncTable(); table = true; }                      // solidcloud.cpp:226
        std::vector<std::unique_ptr<IShape>> shapes;
        std::vector<Solid> solids;
        for (int s = 0; s < n_solids; ++s) {
            char tmpl[] = "/tmp/sdfibm_ref_dict_XXXXXX";
            const int fd = mkstemp(tmpl);
            if (fd < 0) return -3;
            close(fd);
            { std::ofstream os(tmpl); os << dict_text[s] << "\n"; }
            Foam::dictionary d = Foam::dictionary::fromFile(tmpl);
            std::remove(tmpl);
            shapes.push_back(ShapeFactory::create(std::string(d.lookup("type")), d));
            solids.emplace_back(s, Foam::vector(pos[3 * s], pos[3 * s + 1], pos[3 * s + 2]),
                                Foam::quaternion(quat[4 * s], Foam::vector(quat[4 * s + 1], quat[4 * s + 2], quat[4 * s + 3])));
            solids.back().setShape(shapes.back().get());
        }
        BBox bbox(bmin, bmax);
        UGrid grid(bbox, delta);                                                     // solidcloud.cpp:245
        grid.clear();
        for (const Solid &s : solids) {                                             // :479-484
            vector c = s.getCenter();
            grid.insert(c.x(), c.y(), c.z(), s.getID());
        }
        std::vector<CollisionPair> cps;
        grid.generateCollisionPairs(cps);
        int64_t n = 0;
        for (CollisionPair &cp : cps) {
            if (n >= cap) return -1;
            pairs[2 * n] = cp.first; pairs[2 * n + 1] = cp.second; ++n;
            Solid &s1 = solids[cp.first], &s2 = solids[cp.second];                   // solidSolidCollision, :492-519
            collisionFunc cfunc = getCollisionFunc(s1.getShape()->getTypeName(), s2.getShape()->getTypeName());
            if (!cfunc) continue;
            vector cP, cN;
            const scalar cd = cfunc(s1, s2, cP, cN);
            if (cd < 0) continue;
            const vector force = 1e4 * cd * cN;
            for (int k = 0; k < 3; ++k) { ft[6 * cp.first + k] -= force[k]; ft[6 * cp.second + k] += force[k]; }
        }
        return n;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_collide: %s\n", e.what());
        return -2;
    }
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// the whole of interact() (reference src/solidcloud.cpp:361-464) around the reference's compiled classes, on a persistent mesh
// ---------------------------------------------------------------------------------------------------------------------
struct RefCtx {
    Foam::fvMesh mesh;
    std::unique_ptr<GeometricTools> geo;
};

static void fill_mesh(Foam::fvMesh &mesh, const ref_mesh *m) {   // views over the caller's arrays (which must outlive the mesh)
    const Foam::vector *P = reinterpret_cast<const Foam::vector *>(m->points);
    mesh.pts = {P, m->n_points};
    mesh.cc = {reinterpret_cast<const Foam::vector *>(m->cc), m->n_cells};
    mesh.cv = {m->V, m->n_cells};
    mesh.fc = {reinterpret_cast<const Foam::vector *>(m->Cf), m->n_faces};
    mesh.fa = {reinterpret_cast<const Foam::vector *>(m->Sf), m->n_faces};
    mesh.c2c = {m->nb_off, m->nb, m->n_cells};
    mesh.c2p = {m->cp_off, m->cp, m->n_cells};
    mesh.cls = {m->cf_off, m->cf, m->n_cells};
    mesh.fcs = {m->fp_off, m->fp, m->n_faces};
}

static Foam::dictionary dict_of(const char *text) {
    char tmpl[] = "/tmp/sdfibm_ref_dict_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd < 0) throw std::runtime_error("mkstemp failed");
    close(fd);
    { std::ofstream os(tmpl); os << text << "\n"; }
    Foam::dictionary d = Foam::dictionary::fromFile(tmpl);
    std::remove(tmpl);
    return d;
}

extern "C" {

void *ref_create(const ref_mesh *m) {
    try {
        RefCtx *c = new RefCtx();
        fill_mesh(c->mesh, m);
        c->geo.reset(new GeometricTools(c->mesh));
        return c;
    } catch (...) { return nullptr; }
}
void ref_destroy(void *h) { delete static_cast<RefCtx *>(h); }

// Solids [solid_begin, solid_end) of n_solids; fields are reset for all cells, results accumulated for those solids only.
// timed_ms = the reference's own timed region (the solid loop + checkAlpha, :442-451).  Returns the number of list entries.
int64_t ref_interact_full(void *h, int n_solids, const char *const *dict_text, const double *pos, const double *quat, const double *vel,
                          const double *omega, const int32_t *seed, const double *U, double dt, double rhof, int two_d,
                          int solid_begin, int solid_end, int32_t *list_off, int32_t *list_cells, int64_t cap, double *As, double *Fs,
                          double *Ts, double *Ct, double *ft, double *timed_ms) {
    try {
        RefCtx &R = *static_cast<RefCtx *>(h);
        const Foam::fvMesh &mesh = R.mesh;
        GeometricTools &geo = *R.geo;
        const int nC = mesh.nCells();
        std::vector<std::unique_ptr<IShape>> shapes(n_solids);
        std::vector<Solid> solids;
        for (int s = 0; s < n_solids; ++s) {
            if (s >= solid_begin && s < solid_end) {
                Foam::dictionary d = dict_of(dict_text[s]);
                shapes[s] = ShapeFactory::create(std::string(d.lookup("type")), d);
            }
            solids.emplace_back(s, Foam::vector(pos[3 * s], pos[3 * s + 1], pos[3 * s + 2]),
                                Foam::quaternion(quat[4 * s], Foam::vector(quat[4 * s + 1], quat[4 * s + 2], quat[4 * s + 3])));
            solids.back().setShape(shapes[s].get());
            solids.back().setVelocity(Foam::vector(vel[3 * s], vel[3 * s + 1], vel[3 * s + 2]));
            solids.back().setOmega(Foam::vector(omega[3 * s], omega[3 * s + 1], omega[3 * s + 2]));
        }
        // interact(): reset the four fields (:438-441)
        std::vector<Foam::vector> mFs(nC, Foam::vector::zero);
        for (int c = 0; c < nC; ++c) { As[c] = 0.0; Ts[c] = 0.0; Ct[c] = 0.0; }
        for (int i = 0; i < 6 * n_solids; ++i) ft[i] = 0.0;
        if (list_off) { list_off[0] = 0; for (int i = 0; i < 3 * n_solids; ++i) list_off[i + 1] = 0; }
        int64_t n_out = 0;
        const auto t1 = std::chrono::high_resolution_clock::now();
        for (int sid = solid_begin; sid < solid_end; ++sid) {
            // solidFluidInteract (:361-433); the seed is meshSearch::findNearestCell's (OpenFOAM's), supplied by the caller
            Solid &solid = solids[sid];
            CellEnumerator ce(mesh, [&](const vector &v) { return solid.phi01(v); }, seed[sid]);
            auto is = ce.intersect();
            using CT = CellEnumerator::CELL_TYPE;
            size_t num_inside_cells = is[CT::ALL_INSIDE].size();
            std::vector<size_t> cellids;
            cellids.insert(cellids.end(), is[CT::ALL_INSIDE].begin(), is[CT::ALL_INSIDE].end());
            cellids.insert(cellids.end(), is[CT::CENTER_INSIDE].begin(), is[CT::CENTER_INSIDE].end());
            cellids.insert(cellids.end(), is[CT::CENTER_OUTSIDE].begin(), is[CT::CENTER_OUTSIDE].end());
            const int insideType = solid.getID() + 4;
            for (auto cellid : is[CT::ALL_INSIDE]) Ct[cellid] = insideType;
            for (auto cellid : is[CT::CENTER_INSIDE]) Ct[cellid] = CT::CENTER_INSIDE;
            for (auto cellid : is[CT::CENTER_OUTSIDE]) Ct[cellid] = CT::CENTER_OUTSIDE;
            geo.clearCache();
            const scalar dtINV = 1.0 / dt;
            vector force = vector::zero, torque = vector::zero;
            for (size_t counter = 0; counter < cellids.size(); ++counter) {
                const auto cellid = cellids[counter];
                scalar alpha = num_inside_cells > 0 ? 1.0 : 0.0;
                if (counter >= num_inside_cells) alpha = geo.calcCellVolume(cellid, solid, two_d != 0) / mesh.cv[cellid];
                As[cellid] += alpha;
                const vector uf(U[3 * cellid], U[3 * cellid + 1], U[3 * cellid + 2]);
                const vector us = solid.evalPointVelocity(mesh.cc[cellid]);
                const vector f_ = alpha * (uf - us);
                const vector t_ = (mesh.cc[cellid] - solid.getCenter()) ^ f_;
                force += f_ * mesh.cv[cellid] * dtINV;
                torque += t_ * mesh.cv[cellid] * dtINV;
                mFs[cellid] += f_ * dtINV;
                Ts[cellid] += alpha;
            }
            force *= rhof;
            torque *= rhof;
            for (int k = 0; k < 3; ++k) { ft[6 * sid + k] = force[k]; ft[6 * sid + 3 + k] = torque[k]; }
            if (list_cells) {
                const CT types[3] = {CT::ALL_INSIDE, CT::CENTER_INSIDE, CT::CENTER_OUTSIDE};
                for (int t = 0; t < 3; ++t) {
                    for (size_t icell : is[types[t]]) { if (n_out >= cap) return -1; list_cells[n_out++] = (int32_t)icell; }
                    list_off[3 * sid + t + 1] = (int32_t)n_out;
                }
            } else n_out += (int64_t)cellids.size();
        }
        for (int c = 0; c < nC; ++c) As[c] = std::min(As[c], 1.0);   // checkAlpha (:564-570)
        const auto t2 = std::chrono::high_resolution_clock::now();
        if (timed_ms) *timed_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        if (list_off) for (int i = 3 * solid_end; i < 3 * n_solids; ++i) list_off[i + 1] = (int32_t)n_out;
        for (int c = 0; c < nC; ++c) { Fs[3 * c] = mFs[c].x(); Fs[3 * c + 1] = mFs[c].y(); Fs[3 * c + 2] = mFs[c].z(); }
        return n_out;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_interact_full: %s\n", e.what());
        return -2;
    }
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// SolidCloud::evolve (reference src/solidcloud.cpp:466-475,521-562) around the reference's own Solid (src/solid.h: applyForcer,
// addMidFluidForceAndTorque, addAcceleration, move, storeOldForce), MotionFactory / the seven motions (src/libmotion) and
// ForcerFactory / the three forcers (src/libforcer), all compiled unmodified.  The sub-iteration loop below is the caller's,
// restated line for line; collisions as in ref_collide when delta > 0 (HEAD's grid yields no pairs).
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

// motion_text[s] / forcer_text[s]: the solidDict entries of the plugin ("type MotionRotor; period 4; ...") or "" for none.
// fluid_ft: n_steps x n x 6 fluid (force, torque) set before each step (Solid::setFluidForceAndTorque, solidcloud.cpp:424), or
// null.  State arrays are updated in place; ft_out[6n] = total force / torque of the LAST sub-iteration; traj (optional,
// n_steps x n x 13) = pos, quat, vel, omega after every step.
int ref_evolve(int n, const char *const *shape_text, const char *const *motion_text, const char *const *forcer_text, const double *rho,
               double *pos, double *quat, double *vel, double *omega, const double *fluid_ft, const double *gravity, double rhof,
               int n_steps, const double *times, double dt, int n_subiter, const double *bmin, const double *bmax, double delta,
               double *ft_out, double *traj) {
    try {
        static bool table = false;
        if (!table) { InitCollisionFuncTable(); table = true; }
        std::vector<std::unique_ptr<IShape>> shapes;
        std::vector<std::unique_ptr<IMotion>> motions(n);
        std::vector<std::unique_ptr<forcer::IForcer>> forcers(n);
        std::vector<IMaterial> materials;
        materials.reserve(n);
        std::vector<Solid> solids;
        solids.reserve(n);
        for (int s = 0; s < n; ++s) {
            Foam::dictionary d = dict_of(shape_text[s]);
            shapes.push_back(ShapeFactory::create(std::string(d.lookup("type")), d));
            materials.emplace_back(rho[s]);
            solids.emplace_back(s, Foam::vector(pos[3 * s], pos[3 * s + 1], pos[3 * s + 2]),
                                Foam::quaternion(quat[4 * s], Foam::vector(quat[4 * s + 1], quat[4 * s + 2], quat[4 * s + 3])));
            Solid &S = solids.back();
            S.setVelocity(Foam::vector(vel[3 * s], vel[3 * s + 1], vel[3 * s + 2]));
            S.setOmega(Foam::vector(omega[3 * s], omega[3 * s + 1], omega[3 * s + 2]));
            S.setShape(shapes.back().get());                                          // order of solidcloud.cpp:171-189
            S.setMaterial(&materials.back());
            if (motion_text[s] && motion_text[s][0]) {
                Foam::dictionary md = dict_of(motion_text[s]);
                motions[s].reset(MotionFactory::create(std::string(md.lookup("type")), md));
                if (!motions[s]) return -4;
                S.setMotion(motions[s].get());
            }
            if (forcer_text[s] && forcer_text[s][0]) {
                Foam::dictionary fd = dict_of(forcer_text[s]);
                forcers[s] = forcer::ForcerFactory::create(std::string(fd.lookup("type")), fd);
                S.setForcer(forcers[s].get());
            }
        }
        const Foam::vector g(gravity[0], gravity[1], gravity[2]);
        std::unique_ptr<UGrid> grid;
        if (bmin && bmax) { BBox bbox(bmin, bmax); grid.reset(new UGrid(bbox, delta)); }
        std::vector<CollisionPair> cps;
        for (int step = 0; step < n_steps; ++step) {
            scalar time = times[step];
            if (fluid_ft)
                for (int s = 0; s < n; ++s) {
                    const double *f = fluid_ft + ((size_t)step * n + s) * 6;
                    solids[s].setFluidForceAndTorque(Foam::vector(f[0], f[1], f[2]), Foam::vector(f[3], f[4], f[5]));
                }
            const scalar dt_sub = dt / n_subiter;                                     // evolve, :521-562
            for (int i = 0; i < n_subiter; ++i) {
                for (Solid &S : solids) S.clearForceAndTorque();
                for (Solid &S : solids) S.applyForcer(time);
                for (Solid &S : solids) S.addMidFluidForceAndTorque();
                for (Solid &S : solids) {                                             // addMidEnvironment, :466-475
                    const scalar rhos = S.getMaterial()->getRho();
                    const Foam::vector gprime = ((rhos - rhof) / rhos) * g;
                    S.addAcceleration(gprime);
                }
                if (grid) {                                                           // solidSolidInteract, :477-519
                    grid->clear();
                    for (const Solid &S : solids) { const vector c = S.getCenter(); grid->insert(c.x(), c.y(), c.z(), S.getID()); }
                    cps.clear();
                    grid->generateCollisionPairs(cps);
                    for (CollisionPair &cp : cps) {
                        Solid &s1 = solids[cp.first], &s2 = solids[cp.second];
                        collisionFunc cfunc = getCollisionFunc(s1.getShape()->getTypeName(), s2.getShape()->getTypeName());
                        if (!cfunc) continue;
                        vector cP, cN;
                        const scalar cd = cfunc(s1, s2, cP, cN);
                        if (cd < 0) continue;
                        const vector force = 1e4 * cd * cN, torque = vector::zero;
                        s1.addForceAndTorque(-force, -torque);
                        s2.addForceAndTorque(force, torque);
                    }
                }
                for (Solid &S : solids) S.move(time, dt_sub);
            }
            for (Solid &S : solids) S.storeOldForce();
            if (traj)
                for (int s = 0; s < n; ++s) {
                    double *o = traj + ((size_t)step * n + s) * 13;
                    const Solid &S = solids[s];
                    for (int k = 0; k < 3; ++k) { o[k] = S.getCenter()[k]; o[7 + k] = S.getVelocity()[k]; o[10 + k] = S.getOmega()[k]; }
                    o[3] = S.getOrientation().w();
                    for (int k = 0; k < 3; ++k) o[4 + k] = S.getOrientation().v()[k];
                }
        }
        for (int s = 0; s < n; ++s) {
            const Solid &S = solids[s];
            for (int k = 0; k < 3; ++k) {
                pos[3 * s + k] = S.getCenter()[k]; vel[3 * s + k] = S.getVelocity()[k]; omega[3 * s + k] = S.getOmega()[k];
                quat[4 * s + 1 + k] = S.getOrientation().v()[k];
                if (ft_out) { ft_out[6 * s + k] = S.getForce()[k]; ft_out[6 * s + 3 + k] = S.getTorque()[k]; }
            }
            quat[4 * s] = S.getOrientation().w();
        }
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_evolve: %s\n", e.what());
        return -2;
    }
}

} // extern "C"

// SolidCloud::fixInternal (reference src/solidcloud.cpp:288-301) around the reference's Solid::evalPointVelocity.
extern "C" int ref_fix_internal(int n_cells, const double *cc, int n_solids, const double *pos, const double *vel, const double *omega,
                                const double *Ct, double *U) {
    std::vector<Solid> solids;
    for (int s = 0; s < n_solids; ++s) {
        solids.emplace_back(s, Foam::vector(pos[3 * s], pos[3 * s + 1], pos[3 * s + 2]), Foam::quaternion(1.0, Foam::vector::zero));
        solids.back().setVelocity(Foam::vector(vel[3 * s], vel[3 * s + 1], vel[3 * s + 2]));
        solids.back().setOmega(Foam::vector(omega[3 * s], omega[3 * s + 1], omega[3 * s + 2]));
    }
    for (int icell = 0; icell < n_cells; ++icell) {
        if (Ct[icell] >= 4) {
            const label id = (label)(Ct[icell] - 4);
            if (id >= n_solids) return -1;
            const Foam::vector u = solids[id].evalPointVelocity(Foam::vector(cc[3 * icell], cc[3 * icell + 1], cc[3 * icell + 2]));
            U[3 * icell] = u.x(); U[3 * icell + 1] = u.y(); U[3 * icell + 2] = u.z();
        }
    }
    return 0;
}

// SolidCloud::calcMeanField<vector> (reference src/solidcloud.cpp:315-359): the volume-weighted mean of a vector field over a
// substitute shape, with the enumerator driven cell by cell (Empty / GetCurCellInd / GetCurCellType / Next) as the reference does.
// out[4] = mean.xyz, volume.
extern "C" int ref_mean_field(void *h, const char *shape_text, const double *pos, const double *quat, int seed, const double *field,
                              int two_d, double *out) {
    try {
        RefCtx &R = *static_cast<RefCtx *>(h);
        Foam::dictionary d = dict_of(shape_text);
        std::unique_ptr<IShape> shape = ShapeFactory::create(std::string(d.lookup("type")), d);
        Solid tmpSolid(0, Foam::vector(pos[0], pos[1], pos[2]), Foam::quaternion(quat[0], Foam::vector(quat[1], quat[2], quat[3])));
        tmpSolid.setShape(shape.get());
        R.geo->clearCache();
        Foam::vector meanField = Foam::vector::zero;
        scalar volume = 0.0;
        CellEnumerator ce(R.mesh, [&](const vector &v) { return tmpSolid.phi01(v); }, seed);
        scalar alpha = 0.0;
        while (!ce.Empty()) {
            const int icur = ce.GetCurCellInd();
            if (ce.GetCurCellType() == CellEnumerator::CELL_TYPE::ALL_INSIDE) alpha = 1.0;
            else alpha = R.geo->calcCellVolume(icur, tmpSolid, two_d != 0) / R.mesh.cv[icur];
            const scalar dV = alpha * R.mesh.cv[icur];
            volume += dV;
            meanField += dV * Foam::vector(field[3 * icur], field[3 * icur + 1], field[3 * icur + 2]);
            ce.Next();
        }
        const Foam::vector m = meanField / volume;
        out[0] = m.x(); out[1] = m.y(); out[2] = m.z(); out[3] = volume;
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_mean_field: %s\n", e.what());
        return -2;
    }
}

// cloud.out rows (reference src/solidcloud.cpp:595-614 around operator<< / write2D of src/solid.cpp:5-29, compiled unmodified):
// one row per solid, "time " + 18 (3-D) or 9 (2-D) columns, std::scientific at the default precision.
namespace sdfibm { void write2D(std::ostream &os, const Solid &s); }
extern "C" int64_t ref_state_rows(int n, const double *pos, const double *quat, const double *vel, const double *omega, const double *ft,
                                  double time, int two_d, char *buf, int64_t cap) {
    std::ostringstream os;
    os << std::scientific;                                                           // statefile, solidcloud.cpp:235-236
    for (int s = 0; s < n; ++s) {
        Solid S(s, Foam::vector(pos[3 * s], pos[3 * s + 1], pos[3 * s + 2]),
                Foam::quaternion(quat[4 * s], Foam::vector(quat[4 * s + 1], quat[4 * s + 2], quat[4 * s + 3])));
        S.setVelocity(Foam::vector(vel[3 * s], vel[3 * s + 1], vel[3 * s + 2]));
        S.setOmega(Foam::vector(omega[3 * s], omega[3 * s + 1], omega[3 * s + 2]));
        S.setForce(Foam::vector(ft[6 * s], ft[6 * s + 1], ft[6 * s + 2]));
        S.setTorque(Foam::vector(ft[6 * s + 3], ft[6 * s + 4], ft[6 * s + 5]));
        if (two_d) { os << time << ' '; write2D(os, S); os << '\n'; }
        else os << time << ' ' << S << '\n';
    }
    const std::string t = os.str();
    if ((int64_t)t.size() + 1 > cap) return -1;
    std::memcpy(buf, t.c_str(), t.size() + 1);
    return (int64_t)t.size();
}

// Mass properties and point evaluation of ONE shape as the reference's own constructor / members give them (src/libshape/*.h):
// props[12] = volume, volumeINV, radiusB, moi diag (3), moiINV diag (3), com (3); returns finite (1 / 0) or < 0.
// phi01 / phi at n_pts world points for a solid at (pos, quat) go through Solid::phi01 / Solid::phi (src/solid.h:110-119).
extern "C" int ref_shape_props(const char *shape_text, double *props, const double *pos, const double *quat, int n_pts, const double *pts,
                               unsigned char *inside, double *phi) {
    try {
        Foam::dictionary d = dict_of(shape_text);
        std::unique_ptr<IShape> sh = ShapeFactory::create(std::string(d.lookup("type")), d);
        props[0] = sh->m_volume; props[1] = sh->m_volumeINV; props[2] = sh->getRadiusB();
        props[3] = sh->m_moi[0]; props[4] = sh->m_moi[4]; props[5] = sh->m_moi[8];
        props[6] = sh->m_moiINV[0]; props[7] = sh->m_moiINV[4]; props[8] = sh->m_moiINV[8];
        props[9] = sh->m_com.x(); props[10] = sh->m_com.y(); props[11] = sh->m_com.z();
        if (n_pts > 0) {
            Solid S(0, Foam::vector(pos[0], pos[1], pos[2]), Foam::quaternion(quat[0], Foam::vector(quat[1], quat[2], quat[3])));
            S.setShape(sh.get());
            for (int i = 0; i < n_pts; ++i) {
                const Foam::vector p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
                inside[i] = S.phi01(p) ? 1 : 0;
                phi[i] = S.phi(p);
            }
        }
        return sh->finite ? 1 : 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_shape_props: %s\n", e.what());
        return -2;
    }
}
