// ref_bridge.cpp — TEST INFRASTRUCTURE.  A C entry over the reference's OWN CellEnumerator / GeometricTools / shape classes
// (compiled unmodified from /root/reference/src through oracle/refshim): for each solid, the three candidate lists and the
// clipped volume fraction field, following solidFluidInteract (reference src/solidcloud.cpp:361-410) up to alpha.
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

using namespace sdfibm;

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
        mesh.pts.resize(m->n_points);
        for (int i = 0; i < m->n_points; ++i) mesh.pts[i] = Foam::vector(m->points[3 * i], m->points[3 * i + 1], m->points[3 * i + 2]);
        mesh.cc.resize(m->n_cells); mesh.cv.resize(m->n_cells); mesh.c2c.resize(m->n_cells); mesh.c2p.resize(m->n_cells); mesh.cls.resize(m->n_cells);
        for (int c = 0; c < m->n_cells; ++c) {
            mesh.cc[c] = Foam::vector(m->cc[3 * c], m->cc[3 * c + 1], m->cc[3 * c + 2]);
            mesh.cv[c] = m->V[c];
            for (int k = m->nb_off[c]; k < m->nb_off[c + 1]; ++k) mesh.c2c[c].push_back(m->nb[k]);
            for (int k = m->cp_off[c]; k < m->cp_off[c + 1]; ++k) mesh.c2p[c].push_back(m->cp[k]);
            for (int k = m->cf_off[c]; k < m->cf_off[c + 1]; ++k) mesh.cls[c].push_back(m->cf[k]);
        }
        mesh.fc.resize(m->n_faces); mesh.fa.resize(m->n_faces); mesh.fcs.resize(m->n_faces);
        for (int f = 0; f < m->n_faces; ++f) {
            mesh.fc[f] = Foam::vector(m->Cf[3 * f], m->Cf[3 * f + 1], m->Cf[3 * f + 2]);
            mesh.fa[f] = Foam::vector(m->Sf[3 * f], m->Sf[3 * f + 1], m->Sf[3 * f + 2]);
            for (int k = m->fp_off[f]; k < m->fp_off[f + 1]; ++k) mesh.fcs[f].push_back(m->fp[k]);
        }
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
