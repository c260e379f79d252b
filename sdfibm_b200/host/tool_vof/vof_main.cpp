// vof_main.cpp — Foam-free counterpart of the reference's tool_vof/main.cpp:  sdfibm_b200_vof <caseDir> [-name <field>]
// Case directory: solidDict (the tool's flavour: meta / shapes / solids / planes) and runDict with
//   mesh { cells (nx ny nz); origin (x y z); spacing (dx dy dz); }     a single blockMesh-numbered hex block
// Output: <caseDir>/0_<field> (default field name alpha.water, tool_vof/main.cpp:19-23).
#include <cstdio>
#include <cstring>
#include <iostream>

#include "vofcloud.h"

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <caseDir> [-name <field>]\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1];
    std::string field_name = "alpha.water";
    for (int i = 2; i + 1 < argc; ++i)
        if (!std::strcmp(argv[i], "-name")) field_name = argv[i + 1];
    if (argc == 2) std::cout << "* No field name provided, default to " << field_name << "\n";
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
        Foam::fvMesh mesh(view);
        mesh.setTime(0.0);
        mesh.setCaseDir(dir);
        Foam::volScalarField field(field_name, mesh, 0.0);
        VofCloud cloud(dir + "/solidDict", mesh);
        cloud.writeVOF(field);
        sdfibm_mesh_free(st);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "sdfibm_b200_vof: %s\n", e.what());
        return 1;
    }
    return 0;
}
