// foamlite.h — the small slice of OpenFOAM the sdfibm plugin surface is written against, for builds WITHOUT OpenFOAM.
//
// The reference's entity plugins (src/libshape/*.h, src/libmotion/*.h, src/libforcer/*.h), its Solid (src/solid.h) and its
// SolidCloud façade (src/solidcloud.h) use: scalar/label/word, vector, tensor, quaternion (src/types.h:4-16), dictionary
// (lookup / lookupOrDefault / subDict / toc / found / set / write, src/solidcloud.cpp:14-206,616-665) and — SolidCloud
// only — fvMesh's object registry and the vol fields (src/solidcloud.cpp:209-217).  With OpenFOAM present
// (-DSDFIBM_WITH_OPENFOAM, see foam_adapter.H) the real headers are used and this file is empty; without it the same
// host sources compile against the stand-ins below, which restate OpenFOAM 12's published arithmetic
// (VectorI.H, TensorI.H, quaternionI.H) in the same operation order.
#pragma once
#ifndef SDFIBM_WITH_OPENFOAM

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sdfibm_b200.h"

namespace Foam {

using scalar = double;
using label = int32_t;
using word = std::string;

// ---------------------------------------------------------------------------------------------------------------
// vector / tensor (VectorI.H, TensorI.H)
// ---------------------------------------------------------------------------------------------------------------
class vector {
    scalar v_[3];

public:
    vector() : v_{0, 0, 0} {}
    vector(scalar x, scalar y, scalar z) : v_{x, y, z} {}
    scalar x() const { return v_[0]; }
    scalar y() const { return v_[1]; }
    scalar z() const { return v_[2]; }
    scalar &x() { return v_[0]; }
    scalar &y() { return v_[1]; }
    scalar &z() { return v_[2]; }
    scalar operator[](int i) const { return v_[i]; }
    scalar &operator[](int i) { return v_[i]; }
    vector &operator+=(const vector &b) { v_[0] += b.v_[0]; v_[1] += b.v_[1]; v_[2] += b.v_[2]; return *this; }
    vector &operator-=(const vector &b) { v_[0] -= b.v_[0]; v_[1] -= b.v_[1]; v_[2] -= b.v_[2]; return *this; }
    vector &operator*=(scalar s) { v_[0] *= s; v_[1] *= s; v_[2] *= s; return *this; }
    static const vector zero;
};
inline const vector vector::zero = vector(0, 0, 0);
inline vector operator+(const vector &a, const vector &b) { return {a.x() + b.x(), a.y() + b.y(), a.z() + b.z()}; }
inline vector operator-(const vector &a, const vector &b) { return {a.x() - b.x(), a.y() - b.y(), a.z() - b.z()}; }
inline vector operator-(const vector &a) { return {-a.x(), -a.y(), -a.z()}; }
inline vector operator*(scalar s, const vector &a) { return {s * a.x(), s * a.y(), s * a.z()}; }
inline vector operator*(const vector &a, scalar s) { return {a.x() * s, a.y() * s, a.z() * s}; }
inline vector operator/(const vector &a, scalar s) { return {a.x() / s, a.y() / s, a.z() / s}; }
inline scalar operator&(const vector &a, const vector &b) { return a.x() * b.x() + a.y() * b.y() + a.z() * b.z(); }
inline vector operator^(const vector &a, const vector &b) {
    return {a.y() * b.z() - a.z() * b.y(), a.z() * b.x() - a.x() * b.z(), a.x() * b.y() - a.y() * b.x()};
}
inline scalar magSqr(const vector &a) { return a.x() * a.x() + a.y() * a.y() + a.z() * a.z(); }
inline scalar mag(const vector &a) { return std::sqrt(magSqr(a)); }
inline vector cmptMultiply(const vector &a, const vector &b) { return {a.x() * b.x(), a.y() * b.y(), a.z() * b.z()}; }
inline std::ostream &operator<<(std::ostream &os, const vector &v) { return os << '(' << v.x() << ' ' << v.y() << ' ' << v.z() << ')'; }

class tensor {
    scalar v_[9];

public:
    tensor() : v_{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
    tensor(scalar xx, scalar xy, scalar xz, scalar yx, scalar yy, scalar yz, scalar zx, scalar zy, scalar zz)
        : v_{xx, xy, xz, yx, yy, yz, zx, zy, zz} {}
    scalar operator[](int i) const { return v_[i]; }
    scalar &operator[](int i) { return v_[i]; }
    tensor T() const { return {v_[0], v_[3], v_[6], v_[1], v_[4], v_[7], v_[2], v_[5], v_[8]}; }
    static const tensor I;
};
inline const tensor tensor::I = tensor(1, 0, 0, 0, 1, 0, 0, 0, 1);
inline tensor operator&(const tensor &a, const tensor &b) {
    tensor r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    return r;
}
inline vector operator&(const tensor &a, const vector &b) {
    return {a[0] * b.x() + a[1] * b.y() + a[2] * b.z(), a[3] * b.x() + a[4] * b.y() + a[5] * b.z(),
            a[6] * b.x() + a[7] * b.y() + a[8] * b.z()};
}
inline tensor operator/(const tensor &a, scalar s) {
    tensor r;
    for (int i = 0; i < 9; ++i) r[i] = a[i] / s;
    return r;
}
inline scalar det(const tensor &t) {
    return t[0] * t[4] * t[8] + t[1] * t[5] * t[6] + t[2] * t[3] * t[7] - t[0] * t[5] * t[7] - t[1] * t[3] * t[8] - t[2] * t[4] * t[6];
}
inline tensor inv(const tensor &t) {   // cofactor inverse (TensorI.H)
    const scalar d = det(t);
    return tensor(t[4] * t[8] - t[7] * t[5], t[2] * t[7] - t[1] * t[8], t[1] * t[5] - t[2] * t[4],
                  t[6] * t[5] - t[3] * t[8], t[0] * t[8] - t[2] * t[6], t[3] * t[2] - t[0] * t[5],
                  t[3] * t[7] - t[4] * t[6], t[1] * t[6] - t[0] * t[7], t[0] * t[4] - t[3] * t[1]) / d;
}

// ---------------------------------------------------------------------------------------------------------------
// quaternion (quaternionI.H): q = (w, v)
// ---------------------------------------------------------------------------------------------------------------
class quaternion {
    scalar w_;
    vector v_;

public:
    enum rotationSequence { ZYX, ZYZ, ZXY, ZXZ, YXZ, YXY, YZX, YZY, XYZ, XYX, XZY, XZX };
    quaternion() : w_(0), v_() {}
    quaternion(scalar w, const vector &v) : w_(w), v_(v) {}
    explicit quaternion(const vector &v) : w_(0), v_(v) {}
    quaternion(const vector &axis, scalar theta) : w_(std::cos(0.5 * theta)), v_(std::sin(0.5 * theta) * axis) {}
    quaternion(rotationSequence rs, const vector &angles) {
        if (rs != XYZ) throw std::runtime_error("foamlite: only the XYZ rotation sequence is used by sdfibm");
        *this = quaternion(vector(1, 0, 0), angles.x());
        *this *= quaternion(vector(0, 1, 0), angles.y());
        *this *= quaternion(vector(0, 0, 1), angles.z());
    }
    scalar w() const { return w_; }
    const vector &v() const { return v_; }
    quaternion &operator*=(const quaternion &q) {
        const scalar w0 = w_;
        w_ = w_ * q.w_ - (v_ & q.v_);
        v_ = w0 * q.v_ + q.w_ * v_ + (v_ ^ q.v_);
        return *this;
    }
    quaternion &operator+=(const quaternion &q) { w_ += q.w_; v_ += q.v_; return *this; }
    quaternion &operator/=(scalar s) { w_ /= s; v_ = v_ / s; return *this; }
    void normalise() { *this /= std::sqrt(w_ * w_ + magSqr(v_)); }
    quaternion mulq0v(const vector &u) const { return quaternion(-(v_ & u), w_ * u + (v_ ^ u)); }
    vector transform(const vector &u) const;
    tensor R() const {
        const scalar w2 = w_ * w_, x2 = v_.x() * v_.x(), y2 = v_.y() * v_.y(), z2 = v_.z() * v_.z();
        const scalar txy = 2 * v_.x() * v_.y(), twz = 2 * w_ * v_.z(), txz = 2 * v_.x() * v_.z();
        const scalar twy = 2 * w_ * v_.y(), tyz = 2 * v_.y() * v_.z(), twx = 2 * w_ * v_.x();
        return tensor(w2 + x2 - y2 - z2, txy - twz, txz + twy, txy + twz, w2 - x2 + y2 - z2, tyz - twx, txz - twy, tyz + twx,
                      w2 - x2 - y2 + z2);
    }
    vector eulerAngles(rotationSequence rs) const {
        if (rs != XYZ) throw std::runtime_error("foamlite: only the XYZ rotation sequence is used by sdfibm");
        const scalar w2 = w_ * w_, x2 = v_.x() * v_.x(), y2 = v_.y() * v_.y(), z2 = v_.z() * v_.z();
        return vector(std::atan2(2 * (w_ * v_.x() - v_.y() * v_.z()), w2 - x2 - y2 + z2),
                      std::asin(2 * (v_.x() * v_.z() + w_ * v_.y())),
                      std::atan2(2 * (w_ * v_.z() - v_.x() * v_.y()), w2 + x2 - y2 - z2));
    }
    static const quaternion I;
};
inline const quaternion quaternion::I = quaternion(1, vector(0, 0, 0));
inline quaternion operator*(const quaternion &a, const quaternion &b) {
    return quaternion(a.w() * b.w() - (a.v() & b.v()), a.w() * b.v() + b.w() * a.v() + (a.v() ^ b.v()));
}
inline quaternion operator*(scalar s, const quaternion &q) { return quaternion(s * q.w(), s * q.v()); }
inline quaternion operator*(const quaternion &q, scalar s) { return quaternion(q.w() * s, q.v() * s); }
inline quaternion conjugate(const quaternion &q) { return quaternion(q.w(), -q.v()); }
inline vector quaternion::transform(const vector &u) const { return (mulq0v(u) * conjugate(*this)).v(); }

// ---------------------------------------------------------------------------------------------------------------
// dictionary: OpenFOAM dictionary syntax (nested `name { ... }`, `key tokens ;`, // and /* */ comments)
// ---------------------------------------------------------------------------------------------------------------
class ITstream {   // the token stream of one entry, convertible the way the reference uses it
    std::vector<std::string> tok_;
    std::string key_;

public:
    ITstream() = default;
    ITstream(std::string key, std::vector<std::string> t) : tok_(std::move(t)), key_(std::move(key)) {}
    const std::vector<std::string> &tokens() const { return tok_; }
    scalar toScalar(size_t i = 0) const {
        if (i >= tok_.size()) throw std::runtime_error("entry '" + key_ + "': missing value");
        size_t used = 0;
        scalar v = std::stod(tok_[i], &used);
        if (used != tok_[i].size()) throw std::runtime_error("entry '" + key_ + "': not a number: " + tok_[i]);
        return v;
    }
    operator vector() const {
        // "(x y z)" tokenised as ( x y z ); dimensioned forms "name [dims] value" keep only the last token group
        size_t i = 0;
        while (i < tok_.size() && tok_[i] != "(") ++i;
        if (i + 4 >= tok_.size() + 0 || tok_[i + 4] != ")") throw std::runtime_error("entry '" + key_ + "': expected (x y z)");
        return vector(toScalar(i + 1), toScalar(i + 2), toScalar(i + 3));
    }
    explicit operator std::string() const {
        if (tok_.empty()) throw std::runtime_error("entry '" + key_ + "': missing value");
        return tok_[0];
    }
};
inline scalar readScalar(const ITstream &s) { return s.toScalar(s.tokens().size() - 1); }   // last token: also reads `rho [dims] 1`
inline label readLabel(const ITstream &s) { return (label)std::llround(s.toScalar()); }
inline bool readBool(const ITstream &s) {
    const std::string t = std::string(s);
    if (t == "1" || t == "true" || t == "on" || t == "yes" || t == "y" || t == "t") return true;
    if (t == "0" || t == "false" || t == "off" || t == "no" || t == "n" || t == "f" || t == "none") return false;
    throw std::runtime_error("bad bool: " + t);
}

class dictionary {
    struct Entry {
        std::string key;
        bool is_dict = false;
        std::vector<std::string> tokens;
        std::shared_ptr<dictionary> sub;
    };
    std::vector<Entry> entries_;

    const Entry *find(const word &k) const {
        for (const Entry &e : entries_)
            if (e.key == k) return &e;
        return nullptr;
    }
    static std::vector<std::string> tokenize(std::istream &is) {
        std::vector<std::string> out;
        std::string cur;
        auto flush = [&]() { if (!cur.empty()) { out.push_back(cur); cur.clear(); } };
        char ch;
        while (is.get(ch)) {
            if (ch == '/' && is.peek() == '/') { flush(); std::string skip; std::getline(is, skip); continue; }
            if (ch == '/' && is.peek() == '*') {
                flush();
                is.get(ch);
                char prev = 0;
                while (is.get(ch)) { if (prev == '*' && ch == '/') break; prev = ch; }
                continue;
            }
            if (ch == '"') { flush(); std::string s; std::getline(is, s, '"'); out.push_back(s); continue; }
            if (std::isspace((unsigned char)ch)) { flush(); continue; }
            if (ch == '{' || ch == '}' || ch == ';' || ch == '(' || ch == ')') { flush(); out.push_back(std::string(1, ch)); continue; }
            cur.push_back(ch);
        }
        flush();
        return out;
    }
    void parse(const std::vector<std::string> &t, size_t &i, bool top) {
        while (i < t.size()) {
            if (t[i] == "}") { if (top) throw std::runtime_error("dictionary: unexpected }"); ++i; return; }
            if (t[i] == ";") { ++i; continue; }
            Entry e;
            e.key = t[i++];
            if (i < t.size() && t[i] == "{") {
                ++i;
                e.is_dict = true;
                e.sub = std::make_shared<dictionary>();
                e.sub->parse(t, i, false);
            } else {
                while (i < t.size() && t[i] != ";") {
                    if (t[i] == "{" || t[i] == "}") throw std::runtime_error("dictionary: missing ; after entry " + e.key);
                    e.tokens.push_back(t[i++]);
                }
                if (i < t.size()) ++i;
            }
            bool replaced = false;
            for (Entry &old : entries_)
                if (old.key == e.key) { old = e; replaced = true; }   // later definition wins
            if (!replaced) entries_.push_back(std::move(e));
        }
        if (!top) throw std::runtime_error("dictionary: missing }");
    }

public:
    dictionary() = default;
    explicit dictionary(std::istream &is) { size_t i = 0; parse(tokenize(is), i, true); }
    static dictionary fromFile(const std::string &path) {
        std::ifstream f(path);
        if (!f) throw std::runtime_error("cannot open dictionary file " + path);
        return dictionary(f);
    }
    label size() const { return (label)entries_.size(); }
    std::vector<word> toc() const {
        std::vector<word> r;
        for (const Entry &e : entries_) r.push_back(e.key);
        return r;
    }
    bool found(const word &k) const { return find(k) != nullptr; }
    void remove(const word &k) {
        for (size_t i = 0; i < entries_.size(); ++i)
            if (entries_[i].key == k) { entries_.erase(entries_.begin() + i); return; }
    }
    bool isDict(const word &k) const { const Entry *e = find(k); return e && e->is_dict; }
    const dictionary &subDict(const word &k) const {
        const Entry *e = find(k);
        if (!e || !e->is_dict) throw std::runtime_error("dictionary: no sub-dictionary '" + k + "'");
        return *e->sub;
    }
    dictionary &subDict(const word &k) { return const_cast<dictionary &>(static_cast<const dictionary *>(this)->subDict(k)); }
    ITstream lookup(const word &k) const {
        const Entry *e = find(k);
        if (!e || e->is_dict) throw std::runtime_error("dictionary: keyword '" + k + "' is undefined");
        return ITstream(k, e->tokens);
    }
    vector lookupOrDefault(const word &k, const vector &d) const { return found(k) ? vector(lookup(k)) : d; }
    scalar lookupOrDefault(const word &k, scalar d) const { return found(k) ? readScalar(lookup(k)) : d; }
    label lookupOrDefault(const word &k, label d) const { return found(k) ? readLabel(lookup(k)) : d; }
    void setTokens(const word &k, std::vector<std::string> tokens) {
        for (Entry &e : entries_)
            if (e.key == k) { e.is_dict = false; e.sub.reset(); e.tokens = std::move(tokens); return; }
        Entry e;
        e.key = k;
        e.tokens = std::move(tokens);
        entries_.push_back(std::move(e));
    }
    static std::string fmt(scalar x) { char b[40]; std::snprintf(b, sizeof b, "%.17g", x); return b; }
    void set(const word &k, const vector &v) { setTokens(k, {"(", fmt(v.x()), fmt(v.y()), fmt(v.z()), ")"}); }
    void set(const word &k, scalar v) { setTokens(k, {fmt(v)}); }
    void write(std::ostream &os, bool subDict = true, int indent = 0) const {
        const std::string pad(4 * indent, ' ');
        for (const Entry &e : entries_) {
            if (e.is_dict) {
                os << pad << e.key << "\n" << pad << "{\n";
                e.sub->write(os, true, indent + 1);
                os << pad << "}\n";
            } else {
                os << pad << e.key;
                for (size_t i = 0; i < e.tokens.size(); ++i) {
                    const std::string &t = e.tokens[i];
                    if (t == ")") os << ")";
                    else os << ((i > 0 && e.tokens[i - 1] == "(") ? "" : " ") << t;
                }
                os << ";\n";
            }
        }
        (void)subDict;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// fvMesh / vol fields: what SolidCloud touches (registry lookup by name, cell values, mesh arrays, bounds, time)
// ---------------------------------------------------------------------------------------------------------------
class fvMesh;
template <class Type, int NC>
class volField {
    word name_;
    const fvMesh &mesh_;
    std::vector<scalar> data_;   // cell values, NC components interleaved (what the C ABI takes)

public:
    volField(const word &name, const fvMesh &mesh, const Type &init);
    const word &name() const { return name_; }
    const fvMesh &mesh() const { return mesh_; }
    label size() const { return (label)(data_.size() / NC); }
    scalar *data() { return data_.data(); }
    const scalar *data() const { return data_.data(); }
    void correctBoundaryConditions() {}   // no patches in the stand-alone harness
    bool write() const;                   // ASCII internalField into <case>/<time>/<name>
    volField &operator=(const Type &v);
};
using volScalarField = volField<scalar, 1>;
using volVectorField = volField<vector, 3>;

class fvMesh {
    sdfibm_mesh_t view_;
    std::map<word, void *> registry_;
    scalar rho_ = 1.0, time_ = 0.0;
    std::string case_dir_ = ".";

public:
    explicit fvMesh(const sdfibm_mesh_t &view) : view_(view) {}
    const sdfibm_mesh_t &view() const { return view_; }
    label nCells() const { return view_.n_cells; }
    void checkIn(const word &name, void *obj) { registry_[name] = obj; }
    template <class T>
    const T &lookupObject(const word &name) const {
        auto it = registry_.find(name);
        if (it == registry_.end()) throw std::runtime_error("fvMesh: object '" + name + "' is not registered");
        return *static_cast<const T *>(it->second);
    }
    // constant/transportProperties rho (src/solidcloud.cpp:247-249) and runTime
    void setTransportRho(scalar rho) { rho_ = rho; }
    scalar transportRho() const { return rho_; }
    void setTime(scalar t) { time_ = t; }
    scalar timeValue() const { return time_; }
    void setCaseDir(const std::string &d) { case_dir_ = d; }
    const std::string &caseDir() const { return case_dir_; }
};

template <class Type, int NC>
volField<Type, NC>::volField(const word &name, const fvMesh &mesh, const Type &init) : name_(name), mesh_(mesh), data_((size_t)mesh.nCells() * NC) {
    *this = init;
    const_cast<fvMesh &>(mesh).checkIn(name, this);
}
template <>
inline volField<scalar, 1> &volField<scalar, 1>::operator=(const scalar &v) { for (scalar &x : data_) x = v; return *this; }
template <>
inline volField<vector, 3> &volField<vector, 3>::operator=(const vector &v) {
    for (size_t i = 0; i < data_.size(); i += 3) { data_[i] = v.x(); data_[i + 1] = v.y(); data_[i + 2] = v.z(); }
    return *this;
}
template <class Type, int NC>
bool volField<Type, NC>::write() const {
    char tdir[64];
    std::snprintf(tdir, sizeof tdir, "%g", mesh_.timeValue());
    const std::string path = mesh_.caseDir() + "/" + tdir + "_" + name_;
    std::ofstream os(path);
    if (!os) return false;
    os.precision(17);
    os << "internalField nonuniform List<" << (NC == 1 ? "scalar" : "vector") << ">\n" << size() << "\n(\n";
    for (label c = 0; c < size(); ++c) {
        if (NC == 1) os << data_[c] << "\n";
        else os << '(' << data_[3 * (size_t)c] << ' ' << data_[3 * (size_t)c + 1] << ' ' << data_[3 * (size_t)c + 2] << ")\n";
    }
    os << ")\n";
    return true;
}

} // namespace Foam
#endif // !SDFIBM_WITH_OPENFOAM
