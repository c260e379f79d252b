// imaterial.h — density holder (reference src/libmaterial/imaterial.h:8-26).
#pragma once
#include "../types.h"

namespace sdfibm {
class IMaterial {
    scalar m_rho;

public:
    IMaterial(scalar rho = 1.0) : m_rho(rho) {}
    const scalar &getRho() const { return m_rho; }
};
class MaterialDefault : public IMaterial {
public:
    MaterialDefault() : IMaterial() {}
};
} // namespace sdfibm
