// foam_shim.h — TEST INFRASTRUCTURE.  The few OpenFOAM names the reference's hot-path translation units use, so that
// /root/reference/src/{cellenumerator,geometrictools}.cpp and libshape/*.h compile UNMODIFIED, from where they lie, into
// oracle/_ref/ (see oracle/Makefile, target ref).  Arithmetic types (vector / tensor / quaternion / dictionary) come from the
// repository's own Foam-free stand-ins (sdfibm_b200/host/foamlite.h: operation order of OpenFOAM's VectorI.H / quaternionI.H);
// this header adds the list types, forAll and an fvMesh that serves the arrays MeshInfo binds (reference src/meshinfo.h:20-29).
#pragma once
#include <algorithm>
#include <cmath>
#define fvMesh foamlite_fvMesh_not_used_here
#include "../../sdfibm_b200/host/foamlite.h"
#undef fvMesh
#include <vector>

#define forAll(list, i) for (Foam::label i = 0; i < (Foam::label)(list).size(); ++i)

namespace Foam {

// Non-owning views over the caller's flat arrays (the arrays of sdfibm_mesh_t): what the reference indexes as List<T> /
// labelListList.  Nothing is copied, so binding a 16.7 M-cell mesh costs nothing and the timed region contains only the
// reference's own work.
template <class T>
struct UList {
    const T *p = nullptr;
    label n = 0;
    label size() const { return n; }
    const T &operator[](label i) const { return p[i]; }
};
using labelList = UList<label>;
using face = labelList;
using cell = labelList;
struct CsrList {   // labelListList / cellList / faceList: row i is val[off[i] .. off[i+1])
    const label *off = nullptr;
    const label *val = nullptr;
    label n = 0;
    label size() const { return n; }
    labelList operator[](label i) const { return labelList{val + off[i], off[i + 1] - off[i]}; }
};
using labelListList = CsrList;
using cellList = CsrList;
using faceList = CsrList;
static_assert(sizeof(vector) == 3 * sizeof(scalar), "vector must be three packed scalars");
using pointField = UList<vector>;
using vectorField = UList<vector>;
using scalarField = UList<scalar>;

// VectorSpaceI.H of the openfoam.org line (README.md:9 pins OpenFOAM 12): `return vs/mag(vs);` — no guard, so a vanishing vector
// (two solids with the same centre, src/libcollision/collision.cpp:10) gives NaN, as in the oracle and the device code.  The
// openfoam.com line returns Zero below ROOTVSMALL instead; the header cannot be checked here (OpenFOAM is not installed), and the
// case is degenerate either way (no contact normal exists).
inline vector normalised(const vector &v) { return v / mag(v); }

class fvMesh {
public:
    labelListList c2c, c2p;
    pointField pts;
    vectorField cc, fc, fa;
    scalarField cv;
    cellList cls;
    faceList fcs;
    const labelListList &cellCells() const { return c2c; }
    const labelListList &cellPoints() const { return c2p; }
    const pointField &points() const { return pts; }
    const vectorField &cellCentres() const { return cc; }
    const scalarField &V() const { return cv; }
    const vectorField &faceCentres() const { return fc; }
    const vectorField &faceAreas() const { return fa; }
    const cellList &cells() const { return cls; }
    const faceList &faces() const { return fcs; }
    label nCells() const { return cc.size(); }
};

} // namespace Foam
