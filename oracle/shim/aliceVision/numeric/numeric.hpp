// Oracle shim for <aliceVision/numeric/numeric.hpp>. TEST INFRASTRUCTURE ONLY.
// Only the few names the matching/feature headers use (RegionsMatcher.hpp:150 Square,
// PointFeature.hpp Vec2f, Regions.hpp:59 Vec2, IndMatchDecorator.hpp:71-81 Mat).
#pragma once
#include <Eigen/Core>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>
namespace aliceVision {
template <class T> inline T Square(T x) { return x * x; }
struct Vec2 { double v[2]; double operator()(int i) const { return v[i]; } double& operator()(int i) { return v[i]; } };
struct Vec2f {
  float v[2];
  Vec2f() : v{0.f, 0.f} {}
  Vec2f(float x, float y) : v{x, y} {}
  float operator()(int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
  template <class T> Vec2 cast() const { return Vec2{{(double)v[0], (double)v[1]}}; }
};
inline Vec2f operator*(float s, const Vec2f& a) { return Vec2f(s * a.v[0], s * a.v[1]); }
struct Mat {
  struct Col { const double* p; double operator()(int i) const { return p[i]; } };
  std::vector<double> d; long r = 0, c = 0;
  Col col(long j) const { return Col{d.data() + j * r}; }
};
}  // namespace aliceVision
