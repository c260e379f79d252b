// solidcloud.cpp — host façade over the C ABI (see solidcloud.h).  Statement order follows the reference's
// src/solidcloud.cpp so the two read side by side; line numbers in comments refer to it.
#define SDFIBM_REGISTER_BUILTINS
#include "solidcloud.h"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <iomanip>
#include <sstream>
#include <stdexcept>

#include "../../include/sdfibm_b200.h"
#include "entitylibrary.h"
#include "libforcer/forcerfactory.h"
#include "libshape/shapefactory.h"
#ifndef SDFIBM_EXTERNAL_PLUGINS
// the built-in plugin sets register themselves in this translation unit; a build that links its own plugin headers instead
// (e.g. the reference's, tests/test_plugin_surface_cpu.py) defines SDFIBM_EXTERNAL_PLUGINS
#include "libforcer/forcers.h"
#include "libmotion/motions.h"
#else
#include "libmotion/imotion.h"
#endif
#ifdef SDFIBM_WITH_OPENFOAM
#include "IFstream.H"
#endif

namespace sdfibm {

label SolidCloud::N_SUBITER = 20;

namespace {
void check(int rc, const char *what) {
    if (rc != SDFIBM_OK) throw std::runtime_error(std::string(what) + ": " + sdfibm_last_error());
}
#ifndef SDFIBM_WITH_OPENFOAM
// Foam-free field / mesh access; foam_adapter.H provides the same four functions for OpenFOAM objects
inline double *cellData(Foam::volScalarField &f) { return f.data(); }
inline double *cellData(Foam::volVectorField &f) { return f.data(); }
inline const sdfibm_mesh_t &meshView(const Foam::fvMesh &m) { return m.view(); }
inline scalar transportRho(const Foam::fvMesh &m) { return m.transportRho(); }
inline scalar meshTime(const Foam::fvMesh &m) { return m.timeValue(); }
inline std::string casePath(const Foam::fvMesh &m, const std::string &f) { return m.caseDir() + "/" + f; }
inline dictionary readDictionaryFile(const std::string &path) {
    dictionary d = dictionary::fromFile(path);
    d.remove("FoamFile");   // OpenFOAM drops the header entry when reading a dictionary file
    return d;
}
#else
inline dictionary readDictionaryFile(const std::string &path) {
    Foam::IFstream is(path);
    return dictionary(is());
}
#endif
} // namespace

#ifndef SDFIBM_WITH_OPENFOAM
bool SolidCloud::isMaster() const { return m_rank == 0; }
#else
bool SolidCloud::isMaster() const { return Foam::Pstream::master(); }
#endif

void SolidCloud::deviceCommId(void *id128) { check(sdfibm_comm_unique_id(id128), "sdfibm_comm_unique_id"); }
void SolidCloud::initDeviceComm(const void *id128, int rank, int n_ranks) {
    ensureDevice();
    check(sdfibm_comm_init(m_ctx, id128, rank, n_ranks), "sdfibm_comm_init");
    m_rank = rank;
    m_reduce = nullptr;   // the sums come back reduced
}

void SolidCloud::log(const std::string &msg) {
    if (isMaster() && logfile) logfile << msg << std::endl;
}

void SolidCloud::initFromDictionary(const Foam::word &dictfile) {   // :14-206
    dictionary root = readDictionaryFile(dictfile);
    log("Init from " + dictfile);

    const dictionary &meta = root.subDict("meta");
    m_ON_FLUID = Foam::readBool(meta.lookup("on_fluid"));
    m_ON_TWOD = Foam::readBool(meta.lookup("on_twod"));
    if (meta.found("on_meanfield")) {   // :29-34
        m_ON_MEANFIELD = true;
        if (isMaster()) {
            meanFieldFile.open(casePath(m_mesh, "meanfield.out"), std::fstream::out);
            meanFieldFile << std::scientific;
        }
    }
    if (meta.found("sampler")) m_sampler = Foam::word(meta.lookup("sampler"));
    m_gravity = meta.lookup("gravity");
    m_writeFrequency = (unsigned)meta.lookupOrDefault("writeFrequency", (label)1);
    if (meta.found("collision_delta")) m_collisionDelta = Foam::readScalar(meta.lookup("collision_delta"));
    {
        std::ostringstream msg;
        msg << "Summary: " << (m_ON_TWOD ? "2D" : "3D") << ' ' << (m_ON_FLUID ? "FSI" : "DEM (fluid disabled)") << ", g = ("
            << m_gravity[0] << ' ' << m_gravity[1] << ' ' << m_gravity[2] << ").";
        log(msg.str());
    }

    // shapes, forces, motions, materials: any failure is fatal, as in the reference (:146-154) — but reported by
    // exception, never exit(), so an embedding application decides
    m_libshape = EntityLibrary<IShape>(root.subDict("shapes"));
    m_radiusB = -1.0;   // :74-75
    if (root.found("forces")) m_libforcer = EntityLibrary<forcer::IForcer>(root.subDict("forces"));

    const dictionary &motions = root.subDict("motions");
    for (const auto &key : motions.toc()) {
        const dictionary &para = motions.subDict(key);
        const std::string type = Foam::word(para.lookup("type"));
        const std::string name = Foam::word(para.lookup("name"));
        m_libmotion[name] = MotionFactory::create(type, para);
        if (m_libmotion[name] == nullptr) throw std::runtime_error("Unrecognized motion type " + type + '\n');
    }
    const dictionary &materials = root.subDict("materials");
    for (const auto &key : materials.toc()) {
        const dictionary &para = materials.subDict(key);
        const std::string type = Foam::word(para.lookup("type"));
        if (type != "General") throw std::runtime_error("Unrecognizable material parameter!");
        const scalar rho = Foam::readScalar(para.lookup("rho"));
        m_libmat[Foam::word(para.lookup("name"))] = new IMaterial(rho);
    }

    const dictionary &solids = root.subDict("solids");
    const auto names = solids.toc();
    for (size_t i = 0; i < names.size(); ++i) {
        const dictionary &solid = solids.subDict(names[i]);
        const vector pos = solid.lookup("pos");
        if (m_ON_TWOD && pos.z() != 0)
            throw std::runtime_error("Solid must has z=0 in 2D simulation, violated by solid # " + std::to_string(i));
        const size_t row = m_solids.add(pos, solid.lookupOrDefault("euler", vector::zero) * M_PI / 180.0);
        m_solids.v[row] = solid.lookupOrDefault("vel", vector::zero);
        m_solids.w[row] = solid.lookupOrDefault("omega", vector::zero);

        const std::string mot_name = Foam::word(solid.lookup("mot_name"));
        const std::string mat_name = Foam::word(solid.lookup("mat_name"));
        const std::string shp_name = Foam::word(solid.lookup("shp_name"));
        if (mot_name != "free") {
            // the reference's operator[] silently inserts a null motion for an unknown name (:186); report it instead
            const auto it = m_libmotion.find(mot_name);
            if (it == m_libmotion.end()) throw std::runtime_error("Unrecognized motion name " + mot_name);
            m_solids.motion[row] = it->second;
        }
        const auto shp = m_libshape.find(shp_name);
        if (shp == m_libshape.end()) throw std::runtime_error("Unrecognized shape name " + shp_name);
        if (solid.found("for_name")) {
            const std::string for_name = Foam::word(solid.lookup("for_name"));
            const auto f = m_libforcer.find(for_name);
            if (f == m_libforcer.end()) throw std::runtime_error("Unrecognized force name " + for_name);
            m_solids.forcer[row] = f->second.get();
        }
        const auto mat = m_libmat.find(mat_name);
        if (mat == m_libmat.end()) throw std::runtime_error("Unrecognized material name " + mat_name);   // null deref in the reference (:198)
        m_solids.setShapeAndMaterial(row, shp->second.get(), mat->second);
    }
    m_solidDict = root;
}

void SolidCloud::buildShapeTable() {
    // shape table in solidDict order; a registered type with no device tag is a hard error (no CPU fallback)
    const dictionary &shapes = m_solidDict.subDict("shapes");
    std::vector<sdfibm_shape_t> &table = m_shapeTable;
    table.clear();
    m_sdfOps.clear();
    std::map<const IShape *, int> index;
    for (const auto &key : shapes.toc()) {
        const std::string name = Foam::word(shapes.subDict(key).lookup("name"));
        const IShape *sh = m_libshape.at(name).get();
        if (index.count(sh)) continue;
        sdfibm_shape_t rec;
        if (!sh->lowerProgram(rec, m_sdfOps) && !sh->lower(rec))   // a composed shape brings its op program, a primitive its tag record
            throw std::runtime_error("shape type '" + sh->getTypeName() + "' has no device tag: implement IShape::lower() "
                                     "(the coupling path has no CPU fallback)");
        index[sh] = (int)table.size();
        table.push_back(rec);
    }
    if (table.empty()) throw std::runtime_error("solidDict defines no shapes");
    m_shapeIndex.resize(m_solids.size());
    for (size_t i = 0; i < m_solids.size(); ++i) m_shapeIndex[i] = index.at(m_solids.shape[i]);
}

// device context: one per rank, on the GPU given by SDFIBM_DEVICE (default 0).  Created at the first device call; the
// static mesh and the shape table are uploaded exactly once.
void SolidCloud::ensureDevice() {
    if (m_ctx) return;
    int device = 0;
    if (const char *d = std::getenv("SDFIBM_DEVICE")) device = std::atoi(d);
    sdfibm_context *ctx = nullptr;
    check(sdfibm_create(device, &ctx), "sdfibm_create");
    m_ctx = ctx;
    if (const char *k = std::getenv("SDFIBM_CELL_SLOTS")) check(sdfibm_set_cell_slots(m_ctx, std::atoi(k)), "sdfibm_set_cell_slots");
    check(sdfibm_set_mesh(m_ctx, &meshView(m_mesh), m_ON_TWOD ? 1 : 0), "sdfibm_set_mesh");
    if (!m_sdfOps.empty()) check(sdfibm_set_shape_programs(m_ctx, m_sdfOps.data(), (int)m_sdfOps.size()), "sdfibm_set_shape_programs");
    check(sdfibm_set_shapes(m_ctx, m_shapeTable.data(), (int)m_shapeTable.size()), "sdfibm_set_shapes");
}

void SolidCloud::stageRecords() {
    m_solids.pack(m_records, m_shapeIndex);
}

SolidCloud::SolidCloud(const Foam::word &dictfile, Foam::volVectorField &U, scalar time)   // :209-274
    : m_mesh(U.mesh()),
      m_Uf(U),
      m_ct(const_cast<Foam::volScalarField &>(m_mesh.lookupObject<Foam::volScalarField>("Ct"))),
      m_As(const_cast<Foam::volScalarField &>(m_mesh.lookupObject<Foam::volScalarField>("As"))),
      m_Fs(const_cast<Foam::volVectorField &>(m_mesh.lookupObject<Foam::volVectorField>("Fs"))),
      m_Ts(const_cast<Foam::volScalarField &>(m_mesh.lookupObject<Foam::volScalarField>("Ts"))) {
    m_time = time;
    m_solids.reserve(10);
    if (isMaster()) logfile.open(casePath(m_mesh, "cloud.log"), std::fstream::out);

    initFromDictionary(Foam::word(dictfile));
    log("Totally [" + std::to_string(m_solids.size()) + "] solids.");

    if (isMaster()) {
        statefile.open(casePath(m_mesh, "cloud.out"), std::fstream::out);
        statefile << std::scientific;
    }

    buildShapeTable();
    m_forceTorque.assign(6 * m_solids.size(), 0.0);

    m_rhof = transportRho(m_mesh);   // :247-251
    if (!m_ON_FLUID) m_rhof = 0.0;

    m_ON_RESTART = meshTime(m_mesh) > 0;   // :254-258
    if (!m_ON_RESTART) initialCorrect();
    log("END OF INIT");
}

SolidCloud::~SolidCloud() {
    log("Simulation finished! Congratulations!");
    if (m_ctx) sdfibm_destroy(m_ctx);
    for (auto &kv : m_libmotion) delete kv.second;
    for (auto &kv : m_libmat) delete kv.second;
}

void SolidCloud::initialCorrect() {   // :276-286
    this->interact(0, 1);
    m_As.write();
    log("Initial As written to 0 directory");
    m_solids.clearLoads();
}

void SolidCloud::fixInternal(scalar) {   // :288-301 — Ct of the last interact, solid state AFTER evolve
    if (m_solids.empty()) return;
    ensureDevice();
    stageRecords();
    check(sdfibm_fix_internal(m_ctx, m_records.data(), (int)m_records.size(), cellData(m_Uf)), "sdfibm_fix_internal");
    m_Uf.correctBoundaryConditions();
}

void SolidCloud::interact(scalar time, scalar dt) {   // :435-464
    const auto t1 = std::chrono::high_resolution_clock::now();
    // the four fields are rewritten for every cell by the kernels (the reference zeroes them first, :438-441)
    if (!m_solids.empty()) {
        ensureDevice();
        stageRecords();
        check(sdfibm_interact(m_ctx, m_records.data(), (int)m_records.size(), cellData(m_Uf), dt, m_rhof, cellData(m_As), cellData(m_Fs),
                              cellData(m_Ts), cellData(m_ct), m_forceTorque.data()),
              "sdfibm_interact");
        if (m_reduce) m_reduce(m_forceTorque.data(), (int)m_forceTorque.size());   // :427-431, one sum instead of 2N
        m_solids.setFluidLoads(m_forceTorque.data());   // :432
    } else {
        m_ct = 0.0; m_As = 0.0; m_Fs = vector::zero; m_Ts = 0.0;
    }
    const auto t2 = std::chrono::high_resolution_clock::now();
    m_lastInteractMs = std::chrono::duration<double, std::milli>(t2 - t1).count();
    {
        std::ostringstream msg;   // :453-459, the origin of the "interact() ms/step" metric
        msg << "t = " << time << " [FSI took " << m_lastInteractMs << " ms]";
        log(msg.str());
    }
    m_As.correctBoundaryConditions();
    m_Fs.correctBoundaryConditions();
    m_Ts.correctBoundaryConditions();
}

void SolidCloud::addMidEnvironment() { m_solids.addBuoyantWeight(m_gravity, m_rhof); }   // :466-475

void SolidCloud::solidSolidInteract() {   // :477-519 — broad phase, narrow phase and force law on the device
    // HEAD constructs its UGrid with cell size 2*m_radiusB = -2, which has no cells and never yields a pair
    // (SURVEY Q7): nothing to do for delta <= 0.
    if (m_solids.empty() || !(m_collisionDelta > 0)) return;
    ensureDevice();
    stageRecords();
    std::vector<double> ft(6 * m_solids.size(), 0.0);
    int64_t n_pairs = 0;
    check(sdfibm_collide(m_ctx, m_records.data(), (int)m_records.size(), m_collisionDelta, nullptr, 0, &n_pairs, ft.data()), "sdfibm_collide");
    if (n_pairs == 0) return;
    m_solids.addLoads(ft.data());
}

void SolidCloud::evolve(scalar time, scalar dt) {   // :521-562
    m_time = time;
    if (m_solids.size() == 1) N_SUBITER = 1;   // sticky, like the reference's function-static
    const scalar dt_sub = dt / N_SUBITER;
    for (int i = 0; i < N_SUBITER; ++i) {   // five batched passes over the state arrays per sub-iteration
        m_solids.clearLoads();
        m_solids.applyForcers(time);
        m_solids.blendFluidLoads();
        this->addMidEnvironment();
        this->solidSolidInteract();
        m_solids.advance(time, dt_sub);   // `time` is not advanced across sub-iterations (Q8)
    }
    m_solids.rememberLoads();
}

scalar SolidCloud::totalSolidVolume() const {   // :572-576
    const sdfibm_mesh_t &mv = meshView(m_mesh);
    const double *as = cellData(const_cast<Foam::volScalarField &>(m_As));
    scalar sum = 0;
    for (label c = 0; c < mv.n_cells; ++c) sum += as[c] * mv.cell_volumes[c];
    return sum;
}

void SolidCloud::saveState() {   // :578-593
    // Only the master writes files, but the sampler's cross-rank reduction is a collective: every rank takes part in it
    // (the reference calls it from its master-only branch, :580-590, which cannot work in parallel — SURVEY §2).
    if (isMaster() && m_timeStepCounter % m_writeFrequency == 0) {
        statefile << (*this);
        statefile.flush();
    }
    if (m_ON_MEANFIELD) this->writeMeanField();
    ++m_timeStepCounter;
}

// Mean-field sampler (:303-359): for every solid, the average of U over a SUBSTITUTE shape placed at the solid,
// sum(alpha V U) / sum(alpha V) over that shape's candidate cells.  It is the interact path with other inputs: with the
// solids at rest, dt = 1 and rhof = 1 the per-solid "force" of interact IS sum(alpha V U), and with U = (1,0,0) its x
// component is sum(alpha V) — sdfibm_mean_field runs the same kernels twice on scratch outputs.  Unlike the reference this
// does not overwrite Ct as a side effect and, in parallel, reduces on every rank instead of hanging on the master.
void SolidCloud::calcMeanField(std::vector<double> &out) {
    out.assign(3 * m_solids.size(), 0.0);
    if (m_solids.empty() || m_sampler.empty()) return;
    const auto sh = m_libshape.find(m_sampler);
    if (sh == m_libshape.end()) throw std::runtime_error("Unrecognized sampler shape name " + m_sampler);
    int sampler_index = -1;
    {
        const dictionary &shapes = m_solidDict.subDict("shapes");
        int i = 0;
        for (const auto &key : shapes.toc()) {
            if (std::string(Foam::word(shapes.subDict(key).lookup("name"))) == m_sampler) sampler_index = i;
            ++i;
        }
    }
    ensureDevice();
    const size_t n = m_solids.size();
    std::vector<sdfibm_solid_t> recs(n);
    m_solids.pack(recs, {}, sampler_index);
    // this rank's raw sums; numerator and denominator are reduced across ranks BEFORE the division (:353-357): a rank whose
    // block does not touch the solid contributes zeros
    std::vector<double> nd(4 * n), den(n);
    check(sdfibm_mean_field_sums(m_ctx, recs.data(), (int)n, cellData(m_Uf), out.data(), den.data()), "sdfibm_mean_field_sums");
    for (size_t i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) nd[4 * i + d] = out[3 * i + d];
        nd[4 * i + 3] = den[i];
    }
    if (m_reduce) m_reduce(nd.data(), (int)nd.size());
    for (size_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) out[3 * i + d] = nd[4 * i + d] / nd[4 * i + 3];
}

void SolidCloud::writeMeanField() {   // :303-313
    if (m_sampler.empty()) return;
    std::vector<double> mean;
    calcMeanField(mean);
    if (!isMaster()) return;
    for (size_t i = 0; i < m_solids.size(); ++i) meanFieldFile << mean[3 * i] << ' ' << mean[3 * i + 1] << ' ' << mean[3 * i + 2] << ' ';
    meanFieldFile << '\n';
    meanFieldFile.flush();
}

std::ostream &operator<<(std::ostream &os, const SolidCloud &sc) {   // :595-614
    for (size_t i = 0; i < sc.m_solids.size(); ++i) {
        os << sc.m_time << ' ';
        sc.m_solids.writeRow(os, i, sc.m_ON_TWOD);
        os << '\n';
    }
    return os;
}

void SolidCloud::saveRestart(const std::string &filename) {   // :616-665
    dictionary &solids = m_solidDict.subDict("solids");
    const auto names = solids.toc();
    for (size_t i = 0; i < names.size(); ++i) {
        const Solid s(m_solids, i);
        dictionary &solid = solids.subDict(names[i]);
        vector tmp = s.getCenter();
        if (m_ON_TWOD) tmp.z() = 0.0;
        solid.set("pos", tmp);
        tmp = s.getVelocity();
        if (m_ON_TWOD) tmp.z() = 0.0;
        solid.set("vel", tmp);
        solid.set("euler", s.getOrientation().eulerAngles(quaternion::XYZ) * 180.0 / M_PI);   // degrees
        tmp = s.getOmega();
        if (m_ON_TWOD) {
            tmp.x() = 0.0;
            tmp.y() = 0.0;
        }
        solid.set("omega", tmp);
    }
    std::ofstream os(filename);
    if (!os) throw std::runtime_error("cannot write restart dictionary " + filename);
    os << "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       dictionary;\n    object      solidDict;\n}\n";
    m_solidDict.write(os, false);
    const std::time_t now = std::time(nullptr);
    char stamp[64];
    std::strftime(stamp, sizeof stamp, "%Y-%m-%d %H:%M:%S", std::localtime(&now));
    os << "\n// " << stamp << "\n";
}

} // namespace sdfibm
