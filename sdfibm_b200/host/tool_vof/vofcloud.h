// vofcloud.h — the smooth-phase-field initialiser of the reference's tool_vof (tool_vof/solidcloud.{h,cpp}, tool_vof/main.cpp):
// reads a solidDict of the tool's flavour (meta.on_twod, shapes{}, solids{ shp_name pos euler }, planes{ ... }), and writes the
// volume fraction every mesh cell has inside the union of the bodies into the field `name`.  The fractions come from the same
// sm_100a kernels as SolidCloud::interact's As (sdfibm_volume_fraction); the tool's own CPU flood fill is not restated here.
#pragma once
#include <cmath>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/sdfibm_b200.h"
#include "../entitylibrary.h"
#include "../libshape/shapefactory.h"
#include "../solid.h"
#include "../types.h"

namespace sdfibm {

class VofCloud {
    const Foam::fvMesh &m_mesh;
    bool m_ON_TWOD{false};
    EntityLibrary<IShape> m_libshape;
    SolidStates m_solids, m_planes;                   // tool_vof/solidcloud.h: two lists, planes applied after the solids
    std::vector<sdfibm_shape_t> m_shapeTable;
    std::vector<sdfibm_sdf_op_t> m_sdfOps;
    std::vector<sdfibm_solid_t> m_records;
    std::vector<double> m_alpha;
    double m_total{0};
    sdfibm_context *m_ctx{nullptr};

    static void check(int rc, const char *what) {
        if (rc) throw std::runtime_error(std::string(what) + ": " + sdfibm_last_error());
    }
    void readBodies(const dictionary &block, SolidStates &into) {   // tool_vof/solidcloud.cpp:82-105 (and the planes loop)
        const auto names = block.toc();
        for (size_t i = 0; i < names.size(); ++i) {
            const dictionary &d = block.subDict(names[i]);
            const vector pos = d.lookup("pos");
            if (m_ON_TWOD && pos.z() != 0) throw std::runtime_error("Solid must has z=0 in 2D simulation, violated by solid # " + std::to_string(i));
            const size_t row = into.add(pos, d.lookupOrDefault("euler", vector::zero) * M_PI / 180.0);
            const std::string shp_name = Foam::word(d.lookup("shp_name"));
            const auto shp = m_libshape.find(shp_name);
            if (shp == m_libshape.end()) throw std::runtime_error("Unrecognized shape name " + shp_name);
            into.shape[row] = shp->second.get();
        }
    }

public:
    VofCloud(const Foam::word &dictfile, const Foam::fvMesh &mesh) : m_mesh(mesh) {   // tool_vof/solidcloud.cpp:12-114
        dictionary root = dictionary::fromFile(dictfile);
        root.remove("FoamFile");
        m_ON_TWOD = Foam::readBool(root.subDict("meta").lookup("on_twod"));
        m_libshape = EntityLibrary<IShape>(root.subDict("shapes"));
        readBodies(root.subDict("solids"), m_solids);
        if (root.found("planes")) readBodies(root.subDict("planes"), m_planes);
        // shape table in dictionary order + the records: solids first, then planes (writeVOF's two loops, :151-155)
        const dictionary &shapes = root.subDict("shapes");
        std::map<const IShape *, int> index;
        for (const auto &key : shapes.toc()) {
            const IShape *sh = m_libshape.at(Foam::word(shapes.subDict(key).lookup("name"))).get();
            if (index.count(sh)) continue;
            sdfibm_shape_t rec;
            if (!sh->lowerProgram(rec, m_sdfOps) && !sh->lower(rec)) throw std::runtime_error("shape type '" + sh->getTypeName() + "' has no device tag: implement IShape::lower()");
            index[sh] = (int)m_shapeTable.size();
            m_shapeTable.push_back(rec);
        }
        for (const SolidStates *list : {&m_solids, &m_planes})
            for (size_t i = 0; i < list->size(); ++i) {
                sdfibm_solid_t r;
                list->packOne(i, r, index.at(list->shape[i]));
                m_records.push_back(r);
            }
    }
    ~VofCloud() { if (m_ctx) sdfibm_destroy(m_ctx); }
    VofCloud(const VofCloud &) = delete;
    VofCloud &operator=(const VofCloud &) = delete;

    label nSolids() const { return (label)m_solids.size(); }
    label nPlanes() const { return (label)m_planes.size(); }
    const std::vector<double> &alpha() const { return m_alpha; }
    double totalVolume() const { return m_total; }

    // tool_vof/solidcloud.cpp:141-173: fill `field` with the volume fraction, write it, report the total volume
    void writeVOF(Foam::volScalarField &field, std::ostream &info = std::cout) {
        if (!m_ctx) {
            int device = 0;
            if (const char *d = std::getenv("SDFIBM_DEVICE")) device = std::atoi(d);
            check(sdfibm_create(device, &m_ctx), "sdfibm_create");
            check(sdfibm_set_cell_slots(m_ctx, 8), "sdfibm_set_cell_slots");
            check(sdfibm_set_mesh(m_ctx, &m_mesh.view(), m_ON_TWOD ? 1 : 0), "sdfibm_set_mesh");
            if (!m_sdfOps.empty()) check(sdfibm_set_shape_programs(m_ctx, m_sdfOps.data(), (int)m_sdfOps.size()), "sdfibm_set_shape_programs");
            if (!m_shapeTable.empty()) check(sdfibm_set_shapes(m_ctx, m_shapeTable.data(), (int)m_shapeTable.size()), "sdfibm_set_shapes");
        }
        m_alpha.assign((size_t)m_mesh.nCells(), 0.0);
        check(sdfibm_volume_fraction(m_ctx, m_records.data(), (int)m_records.size(), m_alpha.data(), &m_total), "sdfibm_volume_fraction");
        for (size_t c = 0; c < m_alpha.size(); ++c) {
            const double a = m_alpha[c];
            if (a < 0 || a > 1 + 1e-6) info << "Wrong volume fraction " << a << " at cell " << c << '\n';   // :160-161
            field.data()[c] = a;
        }
        field.correctBoundaryConditions();
        field.write();
        info << "* Write smooth phase field [" << field.name() << "] to ./0, total volume = " << m_total << '\n';
    }
};

} // namespace sdfibm
