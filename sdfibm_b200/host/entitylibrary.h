// entitylibrary.h — name -> owned entity map built from one solidDict block (reference src/entitylibrary.h:9-36):
// every sub-dictionary carries `type` (factory key) and `name` (library key).
#pragma once
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>

#include "genericfactory.h"
#include "types.h"

namespace sdfibm {

template <typename BaseType, typename ParaType = dictionary>
class EntityLibrary : public std::unordered_map<std::string, std::unique_ptr<BaseType>> {
public:
    EntityLibrary() = default;
    explicit EntityLibrary(const dictionary &def) {
        const auto keys = def.toc();
        for (const auto &key : keys) {
            const dictionary &d = def.subDict(key);
            const std::string type = Foam::word(d.lookup("type"));
            const std::string name = Foam::word(d.lookup("name"));
            this->emplace(name, GenericFactory<BaseType, ParaType>::create(type, d));
        }
    }
    friend std::ostream &operator<<(std::ostream &os, const EntityLibrary &lib) {
        os << "Objects in the library:\n";
        int i = 1;
        for (const auto &kv : lib) os << '[' << i++ << "] " << kv.first << std::endl;
        return os;
    }
};

} // namespace sdfibm
