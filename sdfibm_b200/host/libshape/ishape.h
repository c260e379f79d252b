// ishape.h — shape plugin interface (reference src/libshape/ishape.h:10-61) plus the one addition the GPU path needs:
// lower(), which writes the POD record of include/sdfibm_b200.h the device tagged union evaluates.
//
// The members a reference shape fills in its constructor (m_radiusB, m_volume, m_volumeINV, m_com, m_moi, m_moiINV,
// finite) and the virtuals it implements (isInside, signedDistance, getTypeName, description) are unchanged.  A shape
// registered without a device tag (lower() left at its default) is refused when the cloud is built: the coupling path
// has no CPU fallback.
#pragma once
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "../types.h"

namespace sdfibm {

#define SHAPETYPENAME(name)                          \
    static std::string typeName() { return name; }   \
    static bool added;                               \
    virtual std::string getTypeName() const { return name; }

class IShape;

template <typename T>
class _shapecreator {
public:
    static std::unique_ptr<IShape> create(const dictionary &para) { return std::make_unique<T>(para); }
};

class IShape {
public:
    struct Transformation {
        vector t;
        quaternion q;
    };
    const static int m_id = -1;
    scalar m_radiusB{0.0}; // radius of bounding sphere
    scalar m_volume{0.0};
    scalar m_volumeINV{0.0};
    vector m_com{vector::zero};
    tensor m_moi{tensor::I}; // in principal frame, diagonal
    tensor m_moiINV{tensor::I};
    bool finite{true};

    SHAPETYPENAME("IShape")

    inline static vector world2local(const vector &p, const Transformation &tr) { return Foam::conjugate(tr.q).transform(p - tr.t); }

    virtual int getShapeID() const { return m_id; }
    virtual scalar getRadiusB() const { return m_radiusB; }

    // host-side evaluation (diagnostics, tests); the per-step path evaluates the lowered record on the device
    bool phi01(const vector &p, const Transformation &tr) const { return isInside(world2local(p, tr)); }
    scalar phi(const vector &p, const Transformation &tr) const { return signedDistance(world2local(p, tr)); }

    virtual std::string description() const = 0;

    // Fill `out` with this shape's device record and return true.  Default: no device tag.
    virtual bool lower(sdfibm_shape_t &out) const {
        (void)out;
        return false;
    }

    // Composed shapes (libshape/sdfshape.h): append this shape's post-fix SDF program to `table` and fill `out` with the
    // SDFIBM_SHAPE_PROGRAM record that points at it.  Default: not a composed shape.  The cloud asks this first, then lower().
    virtual bool lowerProgram(sdfibm_shape_t &out, std::vector<sdfibm_sdf_op_t> &table) const {
        (void)out;
        (void)table;
        return false;
    }

    virtual ~IShape() {}

protected:
    // common part of every record
    void lowerCommon(sdfibm_shape_t &out, int tag) const {
        out = sdfibm_shape_t{};
        out.tag = tag;
        out.finite = finite ? 1 : 0;
        out.radiusB = m_radiusB;
        out.com[0] = m_com.x();
        out.com[1] = m_com.y();
        out.com[2] = m_com.z();
    }

private:
    virtual bool isInside(const vector &p) const = 0;        // local coordinate
    virtual scalar signedDistance(const vector &p) const = 0; // local coordinate
};

} // namespace sdfibm
