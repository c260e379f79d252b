// mesh_host.cpp — Foam-free mesh container for the stand-alone harness.
//
// In the drop-in, OpenFOAM owns these arrays and MeshInfo binds them by reference
// (reference src/meshinfo.h:20-29; mesh.cells()/faces() at src/geometrictools.cpp:61,66).  Here they are
// derived from the polyMesh primitives (points, faces, owner, neighbour) with OpenFOAM's published
// formulas and ordering conventions, so the shipped golden fields reproduce:
//   faceCentres/faceAreas : fan triangulation about the vertex average
//   cellCentres/V         : pyramid decomposition about the mean of the face centres
//   cellPoints()          : ascending point label;  cells(): owned faces then neighbour faces, ascending
//   cellCells()           : ascending internal-face order
#include "../../include/sdfibm_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#ifdef SDFIBM_MESH_STANDALONE
// libsdfibm_mesh.so: the same helpers without the CUDA library around them (hosts that only need a mesh — the CPU reference arm of
// bench.py, the oracle tests — never map libsdfibm_b200.so)
static thread_local std::string g_mesh_error;
void sdfibm_set_error(const std::string &msg) { g_mesh_error = msg; }
extern "C" const char *sdfibm_mesh_last_error(void) { return g_mesh_error.c_str(); }
#else
void sdfibm_set_error(const std::string &msg); // sdfibm_cuda.cu
#endif

struct sdfibm_mesh_storage {
    int32_t n_cells = 0, n_points = 0, n_faces = 0, n_internal = 0;
    std::vector<double> points, cc, V, Cf, Sf;
    std::vector<int32_t> cp_off, cp, cf_off, cf, fp_off, fp, nb_off, nb, owner, neighbour;
    double bmin[3], bmax[3];
};

namespace {

struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double mag(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

void derive(sdfibm_mesh_storage &m) {
    const int64_t nF = m.n_faces, nC = m.n_cells, nI = m.n_internal;
    auto P = [&](int32_t i) { return V3{m.points[3 * (int64_t)i], m.points[3 * (int64_t)i + 1], m.points[3 * (int64_t)i + 2]}; };
    // ---- face centres and area vectors (primitiveMeshFaceCentresAndAreas) ----
    m.Cf.resize(3 * nF);
    m.Sf.resize(3 * nF);
    for (int64_t f = 0; f < nF; ++f) {
        const int32_t *ids = &m.fp[m.fp_off[f]];
        const int n = m.fp_off[f + 1] - m.fp_off[f];
        V3 C, S;
        if (n == 3) {
            C = (1.0 / 3.0) * (P(ids[0]) + P(ids[1]) + P(ids[2]));
            S = 0.5 * cross(P(ids[1]) - P(ids[0]), P(ids[2]) - P(ids[0]));
        } else {
            V3 sumN = {0, 0, 0}, sumAc = {0, 0, 0};
            double sumA = 0.0;
            V3 fC = P(ids[0]);
            for (int k = 1; k < n; ++k) fC = fC + P(ids[k]);
            fC = fC / (double)n;
            for (int k = 0; k < n; ++k) {
                V3 p0 = P(ids[k]), p1 = P(ids[(k + 1) % n]);
                V3 c = p0 + p1 + fC;
                V3 nn = cross(p1 - p0, fC - p0);
                double a = mag(nn);
                sumN = sumN + nn;
                sumA += a;
                sumAc = sumAc + a * c;
            }
            if (sumA < 1e-150) { C = fC; S = {0, 0, 0}; }
            else { C = (1.0 / 3.0) * sumAc / sumA; S = 0.5 * sumN; }
        }
        m.Cf[3 * f] = C.x; m.Cf[3 * f + 1] = C.y; m.Cf[3 * f + 2] = C.z;
        m.Sf[3 * f] = S.x; m.Sf[3 * f + 1] = S.y; m.Sf[3 * f + 2] = S.z;
    }
    auto CF = [&](int64_t f) { return V3{m.Cf[3 * f], m.Cf[3 * f + 1], m.Cf[3 * f + 2]}; };
    auto SF = [&](int64_t f) { return V3{m.Sf[3 * f], m.Sf[3 * f + 1], m.Sf[3 * f + 2]}; };

    // ---- cells(): owned faces ascending, then neighbour faces ascending ----
    m.cf_off.assign(nC + 1, 0);
    for (int64_t f = 0; f < nF; ++f) ++m.cf_off[m.owner[f] + 1];
    for (int64_t f = 0; f < nI; ++f) ++m.cf_off[m.neighbour[f] + 1];
    for (int64_t c = 0; c < nC; ++c) m.cf_off[c + 1] += m.cf_off[c];
    m.cf.resize(m.cf_off[nC]);
    {
        std::vector<int32_t> cur(m.cf_off.begin(), m.cf_off.end() - 1);
        for (int64_t f = 0; f < nF; ++f) m.cf[cur[m.owner[f]]++] = (int32_t)f;
        for (int64_t f = 0; f < nI; ++f) m.cf[cur[m.neighbour[f]]++] = (int32_t)f;
    }
    // ---- cellCells(): per internal face, own gets nei and nei gets own ----
    m.nb_off.assign(nC + 1, 0);
    for (int64_t f = 0; f < nI; ++f) { ++m.nb_off[m.owner[f] + 1]; ++m.nb_off[m.neighbour[f] + 1]; }
    for (int64_t c = 0; c < nC; ++c) m.nb_off[c + 1] += m.nb_off[c];
    m.nb.resize(m.nb_off[nC]);
    {
        std::vector<int32_t> cur(m.nb_off.begin(), m.nb_off.end() - 1);
        for (int64_t f = 0; f < nI; ++f) {
            m.nb[cur[m.owner[f]]++] = m.neighbour[f];
            m.nb[cur[m.neighbour[f]]++] = m.owner[f];
        }
    }
    // ---- cellPoints(): unique point labels of the cell's faces, ascending ----
    m.cp_off.assign(nC + 1, 0);
    m.cp.clear();
    m.cp.reserve(8 * (size_t)nC);
    {
        std::vector<int32_t> tmp;
        for (int64_t c = 0; c < nC; ++c) {
            tmp.clear();
            for (int32_t k = m.cf_off[c]; k < m.cf_off[c + 1]; ++k) {
                int32_t f = m.cf[k];
                tmp.insert(tmp.end(), m.fp.begin() + m.fp_off[f], m.fp.begin() + m.fp_off[f + 1]);
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            m.cp.insert(m.cp.end(), tmp.begin(), tmp.end());
            m.cp_off[c + 1] = (int32_t)m.cp.size();
        }
    }
    // ---- cell centres and volumes (primitiveMeshCellCentresAndVols) ----
    std::vector<double> cEst(3 * nC, 0.0);
    std::vector<int32_t> nCellFaces(nC, 0);
    for (int64_t f = 0; f < nF; ++f) {
        int32_t c = m.owner[f];
        for (int d = 0; d < 3; ++d) cEst[3 * (int64_t)c + d] += m.Cf[3 * f + d];
        ++nCellFaces[c];
    }
    for (int64_t f = 0; f < nI; ++f) {
        int32_t c = m.neighbour[f];
        for (int d = 0; d < 3; ++d) cEst[3 * (int64_t)c + d] += m.Cf[3 * f + d];
        ++nCellFaces[c];
    }
    for (int64_t c = 0; c < nC; ++c)
        for (int d = 0; d < 3; ++d) cEst[3 * c + d] /= nCellFaces[c];
    m.cc.assign(3 * nC, 0.0);
    m.V.assign(nC, 0.0);
    auto CE = [&](int64_t c) { return V3{cEst[3 * c], cEst[3 * c + 1], cEst[3 * c + 2]}; };
    auto accum = [&](int64_t c, double pyr3Vol, V3 pc) {
        m.cc[3 * c] += pyr3Vol * pc.x; m.cc[3 * c + 1] += pyr3Vol * pc.y; m.cc[3 * c + 2] += pyr3Vol * pc.z;
        m.V[c] += pyr3Vol;
    };
    for (int64_t f = 0; f < nF; ++f) {
        int64_t c = m.owner[f];
        double pyr3Vol = dot(SF(f), CF(f) - CE(c));
        V3 pc = (3.0 / 4.0) * CF(f) + (1.0 / 4.0) * CE(c);
        accum(c, pyr3Vol, pc);
    }
    for (int64_t f = 0; f < nI; ++f) {
        int64_t c = m.neighbour[f];
        double pyr3Vol = dot(SF(f), CE(c) - CF(f));
        V3 pc = (3.0 / 4.0) * CF(f) + (1.0 / 4.0) * CE(c);
        accum(c, pyr3Vol, pc);
    }
    for (int64_t c = 0; c < nC; ++c) {
        if (std::fabs(m.V[c]) > 1e-300) for (int d = 0; d < 3; ++d) m.cc[3 * c + d] /= m.V[c];
        else for (int d = 0; d < 3; ++d) m.cc[3 * c + d] = cEst[3 * c + d];
        m.V[c] *= (1.0 / 3.0);
    }
    // ---- bounds ----
    for (int d = 0; d < 3; ++d) { m.bmin[d] = 1e300; m.bmax[d] = -1e300; }
    for (int64_t p = 0; p < m.n_points; ++p)
        for (int d = 0; d < 3; ++d) {
            m.bmin[d] = std::min(m.bmin[d], m.points[3 * p + d]);
            m.bmax[d] = std::max(m.bmax[d], m.points[3 * p + d]);
        }
}

} // namespace

extern "C" {

int sdfibm_mesh_from_polymesh(int32_t n_points, const double *points, int32_t n_faces, const int32_t *face_off,
                              const int32_t *face_pts, const int32_t *owner, int32_t n_internal,
                              const int32_t *neighbour, sdfibm_mesh_storage **out) {
    if (!points || !face_off || !face_pts || !owner || !out || n_internal > n_faces || (n_internal > 0 && !neighbour)) {
        sdfibm_set_error("sdfibm_mesh_from_polymesh: bad argument");
        return SDFIBM_ERR_ARG;
    }
    auto *m = new sdfibm_mesh_storage();
    m->n_points = n_points; m->n_faces = n_faces; m->n_internal = n_internal;
    m->points.assign(points, points + 3 * (int64_t)n_points);
    m->fp_off.assign(face_off, face_off + n_faces + 1);
    m->fp.assign(face_pts, face_pts + face_off[n_faces]);
    m->owner.assign(owner, owner + n_faces);
    m->neighbour.assign(neighbour, neighbour + n_internal);
    int32_t nc = 0;
    for (int32_t f = 0; f < n_faces; ++f) nc = std::max(nc, owner[f] + 1);
    for (int32_t f = 0; f < n_internal; ++f) nc = std::max(nc, neighbour[f] + 1);
    m->n_cells = nc;
    derive(*m);
    *out = m;
    return SDFIBM_OK;
}

/* One hex block in blockMesh numbering (checked against the shipped meshes M1/M2):
 * point = i + (nx+1) j + (nx+1)(ny+1) k, cell = i + nx j + nx ny k, internal faces by owner (x+, y+, z+),
 * then patches left(x-) right(x+) bottom(y-) top(y+) front(z-) back(z+). */
int sdfibm_mesh_hex_block(int32_t nx, int32_t ny, int32_t nz, const double x0[3], const double dx[3],
                          sdfibm_mesh_storage **out) {
    if (nx < 1 || ny < 1 || nz < 1 || !x0 || !dx || !out) {
        sdfibm_set_error("sdfibm_mesh_hex_block: bad argument");
        return SDFIBM_ERR_ARG;
    }
    const int64_t px = nx + 1, py = ny + 1, pz = nz + 1;
    const int64_t nP = px * py * pz, nC = (int64_t)nx * ny * nz;
    const int64_t nI = (int64_t)(nx - 1) * ny * nz + (int64_t)nx * (ny - 1) * nz + (int64_t)nx * ny * (nz - 1);
    const int64_t nF = nI + 2 * ((int64_t)ny * nz + (int64_t)nx * nz + (int64_t)nx * ny);
    if (nP > 2000000000LL || 4 * nF > 2000000000LL || 8 * nC > 2000000000LL) {
        sdfibm_set_error("sdfibm_mesh_hex_block: mesh exceeds int32 labels");
        return SDFIBM_ERR_ARG;
    }
    auto *m = new sdfibm_mesh_storage();
    m->n_points = (int32_t)nP; m->n_faces = (int32_t)nF; m->n_internal = (int32_t)nI; m->n_cells = (int32_t)nC;
    m->points.resize(3 * nP);
    for (int64_t k = 0; k < pz; ++k)
        for (int64_t j = 0; j < py; ++j)
            for (int64_t i = 0; i < px; ++i) {
                int64_t p = i + px * (j + py * k);
                m->points[3 * p] = x0[0] + i * dx[0];
                m->points[3 * p + 1] = x0[1] + j * dx[1];
                m->points[3 * p + 2] = x0[2] + k * dx[2];
            }
    auto pid = [&](int64_t i, int64_t j, int64_t k) { return (int32_t)(i + px * (j + py * k)); };
    auto cid = [&](int64_t i, int64_t j, int64_t k) { return (int32_t)(i + nx * (j + (int64_t)ny * k)); };
    m->fp_off.resize(nF + 1);
    m->fp.resize(4 * nF);
    m->owner.resize(nF);
    m->neighbour.resize(nI);
    int64_t f = 0;
    auto put = [&](int32_t a, int32_t b, int32_t c, int32_t d, int32_t own, int32_t nei) {
        m->fp_off[f] = (int32_t)(4 * f);
        m->fp[4 * f] = a; m->fp[4 * f + 1] = b; m->fp[4 * f + 2] = c; m->fp[4 * f + 3] = d;
        m->owner[f] = own;
        if (nei >= 0) m->neighbour[f] = nei;
        ++f;
    };
    for (int64_t k = 0; k < nz; ++k)
        for (int64_t j = 0; j < ny; ++j)
            for (int64_t i = 0; i < nx; ++i) {
                int32_t c = cid(i, j, k);
                if (i + 1 < nx) put(pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i + 1, j + 1, k + 1), pid(i + 1, j, k + 1), c, cid(i + 1, j, k));
                if (j + 1 < ny) put(pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i + 1, j + 1, k + 1), pid(i + 1, j + 1, k), c, cid(i, j + 1, k));
                if (k + 1 < nz) put(pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j + 1, k + 1), pid(i, j + 1, k + 1), c, cid(i, j, k + 1));
            }
    for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0; j < ny; ++j)   // left  (x-)
        put(pid(0, j, k), pid(0, j, k + 1), pid(0, j + 1, k + 1), pid(0, j + 1, k), cid(0, j, k), -1);
    for (int64_t k = 0; k < nz; ++k) for (int64_t j = 0; j < ny; ++j)   // right (x+)
        put(pid(nx, j, k), pid(nx, j + 1, k), pid(nx, j + 1, k + 1), pid(nx, j, k + 1), cid(nx - 1, j, k), -1);
    for (int64_t i = 0; i < nx; ++i) for (int64_t k = 0; k < nz; ++k)   // bottom (y-)
        put(pid(i, 0, k), pid(i + 1, 0, k), pid(i + 1, 0, k + 1), pid(i, 0, k + 1), cid(i, 0, k), -1);
    for (int64_t i = 0; i < nx; ++i) for (int64_t k = 0; k < nz; ++k)   // top (y+)
        put(pid(i, ny, k), pid(i, ny, k + 1), pid(i + 1, ny, k + 1), pid(i + 1, ny, k), cid(i, ny - 1, k), -1);
    for (int64_t i = 0; i < nx; ++i) for (int64_t j = 0; j < ny; ++j)   // front (z-)
        put(pid(i, j, 0), pid(i, j + 1, 0), pid(i + 1, j + 1, 0), pid(i + 1, j, 0), cid(i, j, 0), -1);
    for (int64_t i = 0; i < nx; ++i) for (int64_t j = 0; j < ny; ++j)   // back (z+)
        put(pid(i, j, nz), pid(i + 1, j, nz), pid(i + 1, j + 1, nz), pid(i, j + 1, nz), cid(i, j, nz - 1), -1);
    m->fp_off[nF] = (int32_t)(4 * nF);
    derive(*m);
    *out = m;
    return SDFIBM_OK;
}

int sdfibm_mesh_view(const sdfibm_mesh_storage *m, sdfibm_mesh_t *v) {
    if (!m || !v) { sdfibm_set_error("sdfibm_mesh_view: null argument"); return SDFIBM_ERR_ARG; }
    v->n_cells = m->n_cells; v->n_points = m->n_points; v->n_faces = m->n_faces; v->n_internal_faces = m->n_internal;
    v->points = m->points.data(); v->cell_centres = m->cc.data(); v->cell_volumes = m->V.data();
    v->face_centres = m->Cf.data(); v->face_areas = m->Sf.data();
    v->cell_points_off = m->cp_off.data(); v->cell_points = m->cp.data();
    v->cell_faces_off = m->cf_off.data(); v->cell_faces = m->cf.data();
    v->face_points_off = m->fp_off.data(); v->face_points = m->fp.data();
    v->cell_cells_off = m->nb_off.data(); v->cell_cells = m->nb.data();
    for (int d = 0; d < 3; ++d) { v->bounds_min[d] = m->bmin[d]; v->bounds_max[d] = m->bmax[d]; }
    return SDFIBM_OK;
}

int sdfibm_mesh_owner_neighbour(const sdfibm_mesh_storage *m, const int32_t **owner, const int32_t **neighbour) {
    if (!m || !owner || !neighbour) { sdfibm_set_error("sdfibm_mesh_owner_neighbour: null argument"); return SDFIBM_ERR_ARG; }
    *owner = m->owner.data();
    *neighbour = m->neighbour.data();
    return SDFIBM_OK;
}

int sdfibm_mesh_free(sdfibm_mesh_storage *m) {
    delete m;
    return SDFIBM_OK;
}

} // extern "C"
