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

template <class T>
class List : public std::vector<T> {
public:
    using std::vector<T>::vector;
    label size() const { return (label)std::vector<T>::size(); }
};
using labelList = List<label>;
using labelListList = List<labelList>;
using pointField = List<vector>;
using vectorField = List<vector>;
using scalarField = List<scalar>;
using face = labelList;
using cell = labelList;
using faceList = List<face>;
using cellList = List<cell>;

// VectorSpaceI.H: vs / mag(vs), zero for a vanishing vector
inline vector normalised(const vector &v) { const scalar m = mag(v); return m > 1e-300 ? v / m : vector::zero; }

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
