// standalone_main.cpp — Foam-free driver of the coupling loop: sdfibm_b200_run <caseDir>
//
// Follows the order of the reference's time loop (src/main.cpp:38-104) with the fluid solve replaced by a prescribed
// velocity field, so the solid side (interact -> U -= Fs dt -> evolve -> saveState -> fixInternal, restart file at the
// end) can be run and compared without OpenFOAM.  Case directory:
//   solidDict      the reference's schema
//   runDict        mesh { cells (nx ny nz); origin (x y z); spacing (dx dy dz); }   single blockMesh-numbered hex block
//                  fluid { rho 1; U (ux uy uz); }                                   uniform initial velocity
//                  time { deltaT 1e-3; nSteps 10; startTime 0; parallel 0; }
// Which solidDict is read follows src/main.cpp:25-36: the case root at time 0, `<startTime>/solidDict` on a restart
// (`processor0/<startTime>/solidDict` with `parallel 1`) when that file exists.
// Outputs: cloud.out, cloud.log, 0_As (initialCorrect), the restart dictionary `solidDict.restart` and, as main.cpp:93-100 does,
// `<endTime>/solidDict` (`processor0/<endTime>/solidDict` with `parallel 1`).
#include <cstdio>
#include <filesystem>
#include <iostream>
#include <memory>

#include "../../include/sdfibm_b200.h"
#include "solidcloud.h"

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <caseDir>\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1];
    try {
        using namespace sdfibm;
        const dictionary run = dictionary::fromFile(dir + "/runDict");
        const dictionary &md = run.subDict("mesh");
        const vector cells = md.lookup("cells"), origin = md.lookup("origin"), spacing = md.lookup("spacing");
        const double x0[3] = {origin.x(), origin.y(), origin.z()}, dx[3] = {spacing.x(), spacing.y(), spacing.z()};
        sdfibm_mesh_storage *st = nullptr;
        if (sdfibm_mesh_hex_block((int)cells.x(), (int)cells.y(), (int)cells.z(), x0, dx, &st)) throw std::runtime_error(sdfibm_last_error());
        sdfibm_mesh_t view;
        sdfibm_mesh_view(st, &view);

        const dictionary &fd = run.subDict("fluid");
        const dictionary &td = run.subDict("time");
        const scalar dt = Foam::readScalar(td.lookup("deltaT"));
        const label n_steps = Foam::readLabel(td.lookup("nSteps"));
        const scalar t0 = td.lookupOrDefault("startTime", 0.0);

        Foam::fvMesh mesh(view);
        mesh.setTransportRho(Foam::readScalar(fd.lookup("rho")));
        mesh.setTime(t0);
        mesh.setCaseDir(dir);
        // createFields.h:19-35
        Foam::volVectorField U("U", mesh, vector(fd.lookup("U")));
        Foam::volScalarField As("As", mesh, 0.0), Ct("Ct", mesh, 0.0), Ts("Ts", mesh, 0.0);
        Foam::volVectorField Fs("Fs", mesh, vector::zero);

        const bool par = td.lookupOrDefault("parallel", (label)0) != 0;
        auto time_name = [](scalar t) { char b[64]; std::snprintf(b, sizeof b, "%g", t); return std::string(b); };
        std::string dictfile = dir + "/solidDict";                                  // main.cpp:25-36
        if (t0 > 0) {
            const std::string at_time = dir + "/" + SolidCloud::startDictName(time_name(t0), t0, par);
            if (std::filesystem::exists(at_time)) dictfile = at_time;
        }
        SolidCloud solidcloud(dictfile, U, t0);             // main.cpp:38
        solidcloud.saveState();                             // :39
        scalar t = t0;
        for (label step = 0; step < n_steps; ++step) {
            t += dt;   // runTime.loop() advances the clock before the body
            mesh.setTime(t);
            solidcloud.interact(t, dt);                     // :66
            if (solidcloud.isOnFluid()) {                   // :68-71  U = U - Fs*dt
                double *u = U.data();
                const double *f = Fs.data();
                for (size_t i = 0; i < 3 * (size_t)view.n_cells; ++i) u[i] = u[i] - f[i] * dt;
                U.correctBoundaryConditions();
            }
            solidcloud.evolve(t, dt);                       // :82
            solidcloud.saveState();                         // :83
            if (solidcloud.isOnFluid()) solidcloud.fixInternal(dt);   // :85-88
        }
        solidcloud.saveRestart(dir + "/solidDict.restart"); // :101
        {
            const std::string at_time = dir + "/" + SolidCloud::restartDictName(time_name(t), par);   // :93-100
            std::filesystem::create_directories(std::filesystem::path(at_time).parent_path());
            solidcloud.saveRestart(at_time);
        }
        std::printf("ran %d steps, %d solids, last interact %.3f ms, solid volume %.9g\n", (int)n_steps, (int)solidcloud.size(),
                    solidcloud.lastInteractMs(), solidcloud.totalSolidVolume());
        sdfibm_mesh_free(st);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "sdfibm_b200_run: %s\n", e.what());
        return 1;
    }
    return 0;
}
