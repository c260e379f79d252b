// adapter_main.cpp — TEST INFRASTRUCTURE: drives the façade compiled with -DSDFIBM_WITH_OPENFOAM (foam_adapter.H + the OpenFOAM
// branches of solidcloud.cpp) on the mock API.  Usage: adapter_main <caseDir> <nx> <ny> <nz> <startTime> <nSteps> [gpu]
// Builds the mock fvMesh from a Foam-free hex block, registers U / As / Fs / Ts / Ct and transportProperties, then runs the loop of
// src/main.cpp:38-101 (evolve only unless `gpu` is given) and prints the solid states.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <unistd.h>

#include "../../sdfibm_b200/host/solidcloud.h"

using namespace Foam;

int main(int argc, char **argv) {
    if (argc < 7) return 2;
    const std::string dir = argv[1];
    const int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]);
    const double t0 = atof(argv[5]);
    const int n_steps = atoi(argv[6]);
    const bool gpu = argc > 7 && !strcmp(argv[7], "gpu");
    if (chdir(dir.c_str())) return 3;   // OpenFOAM runs with the case directory as cwd (casePath() of the adapter relies on it)
    try {
        const double x0[3] = {-2, -2, -2}, dx[3] = {4.0 / nx, 4.0 / ny, 4.0 / nz};
        sdfibm_mesh_storage *st = nullptr;
        if (sdfibm_mesh_hex_block(nx, ny, nz, x0, dx, &st)) throw std::runtime_error(sdfibm_last_error());
        sdfibm_mesh_t v;
        sdfibm_mesh_view(st, &v);
        fvMesh mesh;
        auto fill = [](const int32_t *off, const int32_t *val, int n, auto &ll) {
            ll.resize(n);
            for (int i = 0; i < n; ++i) ll[i].assign(val + off[i], val + off[i + 1]);
        };
        fill(v.cell_points_off, v.cell_points, v.n_cells, mesh.c2p);
        fill(v.cell_cells_off, v.cell_cells, v.n_cells, mesh.c2c);
        fill(v.cell_faces_off, v.cell_faces, v.n_cells, mesh.cls);
        fill(v.face_points_off, v.face_points, v.n_faces, mesh.fcs);
        auto vec = [](const double *p, int n, auto &f) { f.resize(n); for (int i = 0; i < n; ++i) f[i] = vector(p[3 * i], p[3 * i + 1], p[3 * i + 2]); };
        vec(v.points, v.n_points, mesh.pts);
        vec(v.cell_centres, v.n_cells, mesh.cc);
        vec(v.face_centres, v.n_faces, mesh.fc);
        vec(v.face_areas, v.n_faces, mesh.fa);
        mesh.vol.f.assign(v.cell_volumes, v.cell_volumes + v.n_cells);
        mesh.bb.min_ = vector(v.bounds_min[0], v.bounds_min[1], v.bounds_min[2]);
        mesh.bb.max_ = vector(v.bounds_max[0], v.bounds_max[1], v.bounds_max[2]);
        mesh.n_internal = v.n_internal_faces;
        mesh.runTime.setTime(t0);
        dictionary tp;
        tp.set("rho", 1.25);
        IOdictionary transportProperties(tp);
        mesh.checkIn("transportProperties", &transportProperties);
        volVectorField U("U", mesh, vector(0.1, -0.05, 0.02));
        volScalarField As("As", mesh, 0.0), Ct("Ct", mesh, 0.0), Ts("Ts", mesh, 0.0);
        volVectorField Fs("Fs", mesh, vector::zero);

        sdfibm::SolidCloud cloud("solidDict", U, t0);
        cloud.setForceTorqueReducer(sdfibm::foamSumReduce);
        sdfibm::foamInitDeviceComm(cloud);   // (serial here: returns at once; compiled all the same)
        cloud.saveState();
        double t = t0;
        const double dt = 0.01;
        for (int s = 0; s < n_steps; ++s) {
            t += dt;
            mesh.runTime.setTime(t);
            if (gpu) cloud.interact(t, dt);
            cloud.evolve(t, dt);
            cloud.saveState();
            if (gpu && cloud.isOnFluid()) cloud.fixInternal(dt);
        }
        std::printf("solids %d rho %.17g As_sum ", (int)cloud.size(), 1.25);
        double as = 0;
        for (label c = 0; c < As.size(); ++c) as += As[c];
        std::printf("%.17g\n", as);
        for (label i = 0; i < cloud.size(); ++i) {
            sdfibm_solid_t r;
            cloud[i].toRecord(r, 0);
            std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", r.pos[0], r.pos[1], r.pos[2], r.quat[0], r.quat[1], r.quat[2],
                        r.quat[3], r.vel[0], r.vel[1], r.vel[2]);
        }
        sdfibm_mesh_free(st);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "adapter_main: %s\n", e.what());
        return 1;
    }
    return 0;
}
