// iforcer.h — external force/torque generator interface (reference src/libforcer/iforcer.h:6-41): generate(t, x, v, q, omega) ->
// (F, T), blended 1.5 new - 0.5 old by Solid::applyForcer.  A forcer plugin written for the reference (TYPENAME, _creator<T>,
// IForcer::Force) compiles against this header unchanged (tests/test_plugin_surface_cpu.py).
#pragma once
#include <memory>
#include <string>
#include <utility>

#include "../types.h"

namespace sdfibm {
namespace forcer {

#ifndef TYPENAME   // (imotion.h defines the same macro with the same text, as in the reference)
#define TYPENAME(name)                             \
    static std::string typeName() { return name; } \
    static bool added;
#endif
#define FORCERTYPENAME(name) TYPENAME(name)

class IForcer;
template <typename T>
class _creator {
public:
    static std::unique_ptr<IForcer> create(const dictionary &para) { return std::make_unique<T>(para); }
};

class IForcer {
public:
    using Force = std::pair<vector, vector>;
    IForcer() = default;
    virtual ~IForcer() = default;
    virtual Force generate(const scalar &time, const vector &position, const vector &velocity, const quaternion &orientation,
                           const vector &omega) = 0;
    virtual std::string description() const = 0;
};

} // namespace forcer
} // namespace sdfibm
