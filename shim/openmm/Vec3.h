#ifndef SHIM_OPENMM_VEC3_H_
#define SHIM_OPENMM_VEC3_H_
#include <cassert>
#include <iosfwd>
#include <ostream>
namespace OpenMM {
class Vec3 {
public:
    Vec3() { data[0] = data[1] = data[2] = 0.0; }
    Vec3(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    bool operator==(const Vec3& r) const { return data[0] == r[0] && data[1] == r[1] && data[2] == r[2]; }
    bool operator!=(const Vec3& r) const { return !(*this == r); }
    Vec3 operator+() const { return *this; }
    Vec3 operator+(const Vec3& r) const { return Vec3(data[0] + r[0], data[1] + r[1], data[2] + r[2]); }
    Vec3& operator+=(const Vec3& r) { data[0] += r[0]; data[1] += r[1]; data[2] += r[2]; return *this; }
    Vec3 operator-() const { return Vec3(-data[0], -data[1], -data[2]); }
    Vec3 operator-(const Vec3& r) const { return Vec3(data[0] - r[0], data[1] - r[1], data[2] - r[2]); }
    Vec3& operator-=(const Vec3& r) { data[0] -= r[0]; data[1] -= r[1]; data[2] -= r[2]; return *this; }
    Vec3 operator*(double s) const { return Vec3(data[0] * s, data[1] * s, data[2] * s); }
    Vec3& operator*=(double s) { data[0] *= s; data[1] *= s; data[2] *= s; return *this; }
    Vec3 operator/(double s) const { double i = 1.0 / s; return Vec3(data[0] * i, data[1] * i, data[2] * i); }
    Vec3& operator/=(double s) { double i = 1.0 / s; data[0] *= i; data[1] *= i; data[2] *= i; return *this; }
    double dot(const Vec3& r) const { return data[0] * r[0] + data[1] * r[1] + data[2] * r[2]; }
    Vec3 cross(const Vec3& r) const { return Vec3(data[1] * r[2] - data[2] * r[1], data[2] * r[0] - data[0] * r[2], data[0] * r[1] - data[1] * r[0]); }
private:
    double data[3];
};
static inline Vec3 operator*(double s, const Vec3& v) { return v * s; }
template <class CH, class TR>
std::basic_ostream<CH, TR>& operator<<(std::basic_ostream<CH, TR>& o, const Vec3& v) { o << '[' << v[0] << ", " << v[1] << ", " << v[2] << ']'; return o; }
}  // namespace OpenMM
#endif
