// solidcloud.h — sdfibm::SolidCloud, the one class main.cpp talks to (reference src/solidcloud.h:35-123), with the
// per-step coupling done by the sm_100a library behind include/sdfibm_b200.h instead of the CPU flood fill.
//
// Same constructor and per-step members as the reference (main.cpp:38-39,48,66,68,82-83,85-87,101):
//   SolidCloud(dictfile, U, time) · isOnFluid() · interact(t, dt) · evolve(t, dt) · saveState() · fixInternal(dt) ·
//   saveRestart(file); fields As / Fs / Ts / Ct are found by NAME in U.mesh()'s registry, rho in transportProperties.
// Host-side and unchanged in meaning: solidDict parsing, the shape / motion / forcer / material libraries and their
// factories, Solid::move and the AB2 force blending, cloud.out and the restart dictionary.
// On the device: CellEnumerator + GeometricTools + the forcing loop (interact), fixInternal, and the collision step.
#pragma once
#include <fstream>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "entitylibrary.h"
#include "solid.h"
#include "types.h"

#ifdef SDFIBM_WITH_OPENFOAM
#include "foam_adapter.H"
#endif

struct sdfibm_context;

namespace sdfibm {

class SolidCloud {
private:
    bool m_ON_FLUID{true};
    bool m_ON_TWOD{false};
    bool m_ON_RESTART{false};
    bool m_ON_MEANFIELD{false};
    std::string m_sampler;           // name of the substitute shape of the mean-field sampler (meta.sampler)
    unsigned int m_timeStepCounter{0};
    unsigned int m_writeFrequency{1};
    scalar m_time{0};
    SolidStates m_solids;            // struct of arrays: one row per solid (solid.h)
    dictionary m_solidDict;
    scalar m_radiusB{-1.0};          // reference src/solidcloud.cpp:74-75: never computed, so the UGrid cell is 2*(-1)
    scalar m_collisionDelta{-2.0};   // UGrid cell size handed to the collision step (HEAD value: no pairs)
    vector m_gravity;
    scalar m_rhof{0};

    const Foam::fvMesh &m_mesh;
    Foam::volVectorField &m_Uf;
    Foam::volScalarField &m_ct;
    Foam::volScalarField &m_As;
    Foam::volVectorField &m_Fs;
    Foam::volScalarField &m_Ts;

    std::map<std::string, IMotion *> m_libmotion;
    std::map<std::string, IMaterial *> m_libmat;
    EntityLibrary<IShape> m_libshape;
    EntityLibrary<forcer::IForcer> m_libforcer;
    std::ofstream statefile;
    std::ofstream logfile;
    std::ofstream meanFieldFile;

    // device side
    sdfibm_context *m_ctx{nullptr};
    std::vector<sdfibm_shape_t> m_shapeTable;  // lowered shape records, solidDict order
    std::vector<sdfibm_sdf_op_t> m_sdfOps;     // op programs of the composed shapes (SDFIBM_SHAPE_PROGRAM records point into it)
    std::vector<int> m_shapeIndex;             // per solid: row of the shape table
    std::vector<sdfibm_solid_t> m_records;     // staging of the rigid-body records
    std::vector<double> m_forceTorque;         // [6 N] per-solid (F, T) of the last interact
    int m_rank = 0;
    std::function<void(double *, int)> m_reduce; // cross-rank sum of m_forceTorque (Foam::reduce / NCCL); none = serial
    double m_lastInteractMs{0};

    void solidSolidInteract();
    void stageRecords();
    void buildShapeTable();
    void ensureDevice();
    void log(const std::string &msg);

    static label N_SUBITER;   // function-static in the reference (src/solidcloud.cpp:524): shared by every cloud of the process

public:
    SolidCloud(const Foam::word &dictfile, Foam::volVectorField &U, scalar time = 0.0);
    ~SolidCloud();

    void saveState();
    void initFromDictionary(const Foam::word &dictname);
    void saveRestart(const std::string &filename);
    Solid operator[](label i) const { return Solid(m_solids, (size_t)i); }   // a read-only view of row i
    label size() const { return (label)m_solids.size(); }

    void checkAlpha() const {}   // As <= 1 is applied inside the interact kernel (src/solidcloud.cpp:564-570)
    scalar totalSolidVolume() const;
    bool isOnFluid() const { return m_ON_FLUID; }
    bool isOnTwoD() const { return m_ON_TWOD; }

    void evolve(scalar time, scalar dt);
    void interact(scalar time, scalar dt);
    void addMidEnvironment();
    void fixInternal(scalar dt);
    void initialCorrect();
    void writeMeanField();
    // mean of U over each solid's sampler shape, Σ αV·U / Σ αV (src/solidcloud.cpp:315-359); out[3 N]
    void calcMeanField(std::vector<double> &out);

    // ---- additions of this implementation ----
    // cross-rank sum of the per-solid (F, T) array, replacing the 2N Foam::reduce calls (src/solidcloud.cpp:427-431)
    void setForceTorqueReducer(std::function<void(double *, int)> r) { m_reduce = std::move(r); }
    // The same exchange INSIDE the library (NCCL over NVLink, one ncclAllReduce of 6N fp64 on the context stream right behind the
    // kernels): rank 0 draws the 128-byte id (deviceCommId), the host broadcasts it (Pstream / MPI / a file), every rank calls
    // initDeviceComm.  interact() then returns force / torque already summed; no reducer is needed.
    static void deviceCommId(void *id128);
    void initDeviceComm(const void *id128, int rank, int n_ranks);
    // Which solidDict main.cpp reads and where it writes the restart file (src/main.cpp:25-36,93-100): the case root at time 0,
    // `<time>/solidDict` on a serial restart, `processor0/<time>/solidDict` in a parallel run.
    static std::string startDictName(const std::string &time_name, scalar time_value, bool par_run) {
        if (!(time_value > 0)) return "solidDict";
        return (par_run ? "processor0/" : "") + time_name + "/solidDict";
    }
    static std::string restartDictName(const std::string &time_name, bool par_run) {
        return (par_run ? "./processor0/" : "./") + time_name + "/solidDict";
    }
    // Foam-free parallel hosts: this process's rank (0 = master: the only one that writes cloud.out / meanfield.out).  With
    // OpenFOAM the answer comes from Pstream::master().
    void setRank(int rank) { m_rank = rank; }
    bool isMaster() const;
    // UGrid cell size of the collision broad phase; the reference's HEAD value is 2*m_radiusB = -2 (no pairs, ever).
    // Also read from the optional `meta { collision_delta ...; }` key of solidDict.
    void setCollisionDelta(scalar delta) { m_collisionDelta = delta; }
    double lastInteractMs() const { return m_lastInteractMs; }
    const std::vector<double> &forceTorque() const { return m_forceTorque; }
    static void resetSubIterations() { N_SUBITER = 20; }   // for harnesses that build several clouds in one process

    friend std::ostream &operator<<(std::ostream &os, const SolidCloud &sc);

    SolidCloud(const SolidCloud &) = delete;
    SolidCloud &operator=(const SolidCloud &) = delete;
};

} // namespace sdfibm
