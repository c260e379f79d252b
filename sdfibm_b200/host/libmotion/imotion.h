// imotion.h — motion-constraint plugin interface and factory (reference src/libmotion/imotion.h:7-38,
// motionfactory.{h:11-36,cpp:7-31}).  A motion rewrites (velocity, omega) after the free update in Solid::move.
// Naming convention of the reference: 0 frozen, 1 free, 2 user specified, per (vx vy vz wx wy wz).
#pragma once
#include <iostream>
#include <map>
#include <string>

#include "../types.h"

namespace sdfibm {

#define TYPENAME(name)                             \
    static std::string typeName() { return name; } \
    static bool added;

class IMotion;
template <typename T>
class _creator {
public:
    static IMotion *create(const dictionary &node) { return new T(node); }
};

class IMotion {
public:
    IMotion() = default;
    virtual ~IMotion() = default;
    virtual void constraint(const scalar &time, vector &velocity, vector &omega) = 0;
    virtual std::string description() const = 0;
};

class MotionFactory {
public:
    using TCreateMethod = IMotion *(*)(const dictionary &);
    MotionFactory() = delete;
    static bool add(const std::string &name, TCreateMethod create_method) { return methods().emplace(name, create_method).second; }
    // nullptr for an unknown type: the caller reports it (reference src/solidcloud.cpp:116-118)
    static IMotion *create(const std::string &name, const dictionary &node) {
        const auto it = methods().find(name);
        return it == methods().end() ? nullptr : it->second(node);
    }
    static bool has(const std::string &name) { return methods().count(name) != 0; }
    static void report(std::ostream &os = std::cout) {
        int i = 1;
        for (const auto &kv : methods()) os << '[' << i++ << "] " << kv.first << std::endl;
    }

private:
    static std::map<std::string, TCreateMethod> &methods() {
        static std::map<std::string, TCreateMethod> m;
        return m;
    }
};

#define REGISTERMOTION(m) bool sdfibm::m::added = sdfibm::MotionFactory::add(sdfibm::m::typeName(), sdfibm::m::create);

} // namespace sdfibm
