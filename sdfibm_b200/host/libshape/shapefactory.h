// shapefactory.h — ShapeFactory + REGISTERSHAPE (reference src/libshape/shapefactory.h:9-39).  Including this header in
// exactly one translation unit registers the nine built-in shapes; user plugins add their own REGISTERSHAPE lines.
#pragma once
#include "../genericfactory.h"
#include "../types.h"
#include "ishape.h"
#ifndef SDFIBM_EXTERNAL_PLUGINS
#include "shapes.h"   // the nine built-in shapes (a build that brings its own plugin set defines SDFIBM_EXTERNAL_PLUGINS)
#endif

namespace sdfibm {
MAKESPECIALFACTORY(Shape, IShape, dictionary);

#define REGISTERSHAPE(m) bool sdfibm::m::added = sdfibm::ShapeFactory::add(sdfibm::m::typeName(), sdfibm::m::create);
} // namespace sdfibm

#if defined(SDFIBM_REGISTER_BUILTINS) && !defined(SDFIBM_EXTERNAL_PLUGINS)
REGISTERSHAPE(Circle)
REGISTERSHAPE(Sphere)
REGISTERSHAPE(Ellipse)
REGISTERSHAPE(Ellipsoid)
REGISTERSHAPE(Rectangle)
REGISTERSHAPE(Circle_Tail)
REGISTERSHAPE(Circle_TwoTail)
REGISTERSHAPE(Box)
REGISTERSHAPE(Plane)
#endif
