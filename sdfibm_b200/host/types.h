// types.h — the aliases the plugin sources use (reference src/types.h:9-21).
#pragma once
#ifdef SDFIBM_WITH_OPENFOAM
#include "dictionary.H"
#include "quaternion.H"
#include "tensor.H"
#include "vector.H"
#else
#include "foamlite.h"
#endif

namespace sdfibm {
using Foam::dictionary;
using Foam::label;
using Foam::quaternion;
using Foam::scalar;
using Foam::tensor;
using Foam::vector;
const scalar SMALL = 1e-6;   // reference src/types.h:19
} // namespace sdfibm
