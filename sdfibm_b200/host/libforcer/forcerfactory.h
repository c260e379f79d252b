// forcerfactory.h — ForcerFactory + REGISTERFORCE (reference src/libforcer/forcerfactory.h:6-18).
#pragma once
#include "../genericfactory.h"
#include "../types.h"
#include "iforcer.h"

namespace sdfibm {
namespace forcer {
MAKESPECIALFACTORY(Forcer, IForcer, dictionary);
} // namespace forcer
} // namespace sdfibm
#define REGISTERFORCE(m) bool sdfibm::m::added = sdfibm::forcer::ForcerFactory::add(sdfibm::m::typeName(), sdfibm::m::create);
