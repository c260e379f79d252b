// capi_host.cpp — C binding of the host façade (include/sdfibm_b200_host.h).
#include "../../include/sdfibm_b200_host.h"

#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "libforcer/forcerfactory.h"
#ifndef SDFIBM_EXTERNAL_PLUGINS
#include "libmotion/motions.h"
#endif
#include "libshape/shapefactory.h"
#include "solidcloud.h"
#include "tool_vof/vofcloud.h"

using namespace sdfibm;

static thread_local std::string g_err;
#define HOST_TRY(...)                                    \
    try {                                                \
        __VA_ARGS__;                                     \
        return 0;                                        \
    } catch (const std::exception &e) {                  \
        g_err = e.what();                                \
        return 1;                                        \
    }

struct sdfibm_host_cloud {
    std::unique_ptr<Foam::fvMesh> mesh;
    std::unique_ptr<Foam::volVectorField> U, Fs;
    std::unique_ptr<Foam::volScalarField> As, Ts, Ct;
    std::unique_ptr<SolidCloud> cloud;
};

namespace sdfibm {
// a registered shape type that never got a device tag
class TestNoDeviceTag : public IShape, public _shapecreator<TestNoDeviceTag> {
public:
    TestNoDeviceTag(const dictionary &) {
        m_volume = 1.0;
        m_volumeINV = 1.0;
    }
    SHAPETYPENAME("TestNoDeviceTag")
    virtual std::string description() const override { return "test shape without lower()"; }

private:
    virtual bool isInside(const vector &) const override { return false; }
    virtual scalar signedDistance(const vector &) const override { return 1.0; }
};
bool TestNoDeviceTag::added = false;
} // namespace sdfibm

extern "C" {

const char *sdfibm_host_last_error(void) { return g_err.c_str(); }

int sdfibm_host_create(const char *dictfile, const char *case_dir, const sdfibm_mesh_t *mesh, double rho_fluid, double start_time,
                       const double *U_init, sdfibm_host_cloud **out) {
    HOST_TRY({
        if (!dictfile || !case_dir || !mesh || !out) throw std::runtime_error("sdfibm_host_create: null argument");
        auto h = std::make_unique<sdfibm_host_cloud>();
        h->mesh = std::make_unique<Foam::fvMesh>(*mesh);
        h->mesh->setTransportRho(rho_fluid);
        h->mesh->setTime(start_time);
        h->mesh->setCaseDir(case_dir);
        // createFields.h:19-35 — the names are the contract
        h->U = std::make_unique<Foam::volVectorField>("U", *h->mesh, Foam::vector::zero);
        h->As = std::make_unique<Foam::volScalarField>("As", *h->mesh, 0.0);
        h->Ct = std::make_unique<Foam::volScalarField>("Ct", *h->mesh, 0.0);
        h->Fs = std::make_unique<Foam::volVectorField>("Fs", *h->mesh, Foam::vector::zero);
        h->Ts = std::make_unique<Foam::volScalarField>("Ts", *h->mesh, 0.0);
        if (U_init) std::memcpy(h->U->data(), U_init, sizeof(double) * 3 * (size_t)mesh->n_cells);
        h->cloud = std::make_unique<SolidCloud>(Foam::word(dictfile), *h->U, start_time);
        *out = h.release();
    })
}

int sdfibm_host_destroy(sdfibm_host_cloud *h) {
    delete h;
    return 0;
}

int sdfibm_host_field(sdfibm_host_cloud *h, const char *name, double **data, int64_t *n) {
    HOST_TRY({
        const std::string k(name);
        const int64_t nC = h->mesh->nCells();
        if (k == "U") { *data = h->U->data(); *n = 3 * nC; }
        else if (k == "Fs") { *data = h->Fs->data(); *n = 3 * nC; }
        else if (k == "As") { *data = h->As->data(); *n = nC; }
        else if (k == "Ts") { *data = h->Ts->data(); *n = nC; }
        else if (k == "Ct") { *data = h->Ct->data(); *n = nC; }
        else throw std::runtime_error("unknown field " + k);
    })
}

int sdfibm_host_is_on_fluid(sdfibm_host_cloud *h, int *on_fluid, int *on_twod) {
    HOST_TRY({ *on_fluid = h->cloud->isOnFluid(); *on_twod = h->cloud->isOnTwoD(); })
}
int sdfibm_host_interact(sdfibm_host_cloud *h, double time, double dt) { HOST_TRY(h->cloud->interact(time, dt)) }
int sdfibm_host_evolve(sdfibm_host_cloud *h, double time, double dt) { HOST_TRY(h->cloud->evolve(time, dt)) }
int sdfibm_host_save_state(sdfibm_host_cloud *h) { HOST_TRY(h->cloud->saveState()) }
int sdfibm_host_fix_internal(sdfibm_host_cloud *h, double dt) { HOST_TRY(h->cloud->fixInternal(dt)) }
int sdfibm_host_save_restart(sdfibm_host_cloud *h, const char *filename) { HOST_TRY(h->cloud->saveRestart(filename)) }

// tool_vof: VofCloud(dictfile, mesh).writeVOF(name) on a Foam-free mesh
int sdfibm_host_write_vof(const char *dictfile, const char *case_dir, const sdfibm_mesh_t *mesh, const char *field_name, double *alpha,
                          double *total_volume, int *n_solids, int *n_planes) {
    HOST_TRY({
        if (!dictfile || !case_dir || !mesh || !field_name) throw std::runtime_error("sdfibm_host_write_vof: null argument");
        Foam::fvMesh m(*mesh);
        m.setTime(0.0);
        m.setCaseDir(case_dir);
        Foam::volScalarField field(field_name, m, 0.0);
        VofCloud cloud(Foam::word(dictfile), m);
        std::ostringstream info;
        cloud.writeVOF(field, info);
        if (alpha) std::memcpy(alpha, cloud.alpha().data(), sizeof(double) * cloud.alpha().size());
        if (total_volume) *total_volume = cloud.totalVolume();
        if (n_solids) *n_solids = cloud.nSolids();
        if (n_planes) *n_planes = cloud.nPlanes();
    })
}

int sdfibm_host_n_solids(sdfibm_host_cloud *h, int *n) { HOST_TRY(*n = h->cloud->size()) }
int sdfibm_host_get_solids(sdfibm_host_cloud *h, sdfibm_solid_t *out) {
    HOST_TRY({
        for (label i = 0; i < h->cloud->size(); ++i) (*h->cloud)[i].toRecord(out[i], -1);
    })
}
int sdfibm_host_get_forces(sdfibm_host_cloud *h, double *ft, double *fluid_ft) {
    HOST_TRY({
        for (label i = 0; i < h->cloud->size(); ++i) {
            const Solid &s = (*h->cloud)[i];
            if (ft) {
                for (int d = 0; d < 3; ++d) { ft[6 * i + d] = s.getForce()[d]; ft[6 * i + 3 + d] = s.getTorque()[d]; }
            }
            if (fluid_ft) {
                for (int d = 0; d < 3; ++d) { fluid_ft[6 * i + d] = s.getFluidForce()[d]; fluid_ft[6 * i + 3 + d] = s.getFluidTorque()[d]; }
            }
        }
    })
}
int sdfibm_host_get_masses(sdfibm_host_cloud *h, double *mass) {
    HOST_TRY({ for (label i = 0; i < h->cloud->size(); ++i) mass[i] = (*h->cloud)[i].getMass(); })
}
int sdfibm_host_mean_field(sdfibm_host_cloud *h, double *mean) {
    HOST_TRY({
        std::vector<double> m;
        h->cloud->calcMeanField(m);
        std::memcpy(mean, m.data(), sizeof(double) * m.size());
    })
}
int sdfibm_host_set_collision_delta(sdfibm_host_cloud *h, double delta) { HOST_TRY(h->cloud->setCollisionDelta(delta)) }
int sdfibm_host_reset_subiterations(void) { HOST_TRY(SolidCloud::resetSubIterations()) }

int sdfibm_host_factory_has(const char *kind, const char *type_name, int *found) {
    HOST_TRY({
        const std::string k(kind);
        // (no trial construction: a plugin written for the reference may std::exit() on a dictionary it does not like)
        if (k == "shape") *found = ShapeFactory::has(type_name);
        else if (k == "forcer") *found = forcer::ForcerFactory::has(type_name);
        else if (k == "motion") { *found = MotionFactory::has(type_name);
        } else throw std::runtime_error("unknown factory kind " + k);
    })
}

int sdfibm_host_register_untagged_shape(void) {
    HOST_TRY({ TestNoDeviceTag::added = ShapeFactory::add(TestNoDeviceTag::typeName(), TestNoDeviceTag::create) || TestNoDeviceTag::added; })
}

static std::unique_ptr<IShape> build_shape(const char *dictfile, const char *shape_name) {
    dictionary root = dictionary::fromFile(dictfile);
    const dictionary &shapes = root.subDict("shapes");
    for (const auto &key : shapes.toc()) {
        const dictionary &d = shapes.subDict(key);
        if (std::string(Foam::word(d.lookup("name"))) == shape_name) return ShapeFactory::create(Foam::word(d.lookup("type")), d);
    }
    throw std::runtime_error(std::string("no shape named ") + shape_name);
}

int sdfibm_host_shape_record(const char *dictfile, const char *shape_name, sdfibm_shape_t *record, double props[6]) {
    HOST_TRY({
        auto sh = build_shape(dictfile, shape_name);
        if (!sh->lower(*record)) throw std::runtime_error("shape has no device tag");
        props[0] = sh->m_volume; props[1] = sh->m_volumeINV; props[2] = sh->m_radiusB;
        props[3] = sh->m_moi[0]; props[4] = sh->m_moi[4]; props[5] = sh->m_moi[8];
    })
}

int sdfibm_host_shape_eval(const char *dictfile, const char *shape_name, const double pos[3], const double quat[4], const double *points,
                           int64_t n, int32_t *inside, double *phi) {
    HOST_TRY({
        auto sh = build_shape(dictfile, shape_name);
        const IShape::Transformation tr{vector(pos[0], pos[1], pos[2]), quaternion(quat[0], vector(quat[1], quat[2], quat[3]))};
        for (int64_t i = 0; i < n; ++i) {
            const vector p(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
            if (inside) inside[i] = sh->phi01(p, tr) ? 1 : 0;
            if (phi) phi[i] = sh->phi(p, tr);
        }
    })
}

} // extern "C"
