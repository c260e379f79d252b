// genericfactory.h — name -> creator registry the entity plugins self-register into.
// Same public surface as the reference's (src/genericfactory.h:11-66): GenericFactory<Base, Para>::add / create / report
// and the MAKESPECIALFACTORY macro, so REGISTERSHAPE / REGISTERFORCE plugins written for the reference keep compiling.
#pragma once
#include <functional>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace sdfibm {

template <typename BaseType, typename ParaType>
class GenericFactory {
public:
    using Creator = std::function<std::unique_ptr<BaseType>(const ParaType &)>;
    using CreatorMap = std::unordered_map<std::string, Creator>;

    // false when the name is already taken (the first registration wins, as in the reference)
    static bool add(const std::string &name, Creator fn) { return registry().emplace(name, std::move(fn)).second; }

    static std::unique_ptr<BaseType> create(const std::string &name, const ParaType &para) {
        const auto it = registry().find(name);
        if (it == registry().end()) throw std::runtime_error("Cannot create unrecognized object: " + name + '\n');
        return it->second(para);
    }

    // (an addition: is a creator registered under this name?  A plugin's constructor need not run to find out.)
    static bool has(const std::string &name) { return registry().count(name) != 0; }

    static void report(std::ostream &os = std::cout) {
        os << "Objects the factory can \"produce\":\n";
        int i = 1;
        for (const auto &kv : registry()) os << '[' << i++ << "] " << kv.first << std::endl;
    }

    friend std::ostream &operator<<(std::ostream &os, const GenericFactory &) {
        report(os);
        return os;
    }

private:
    // function-local static: usable from other translation units' static initialisers in any order
    static CreatorMap &registry() {
        static CreatorMap creators;
        return creators;
    }
};

#define MAKESPECIALFACTORY(type, basetype, paratype) using type##Factory = GenericFactory<basetype, paratype>;

} // namespace sdfibm
