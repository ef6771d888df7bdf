// ref_geometry.cpp -- C entry points onto the reference's own, unmodified utils/Geometry.cpp and utils/Utils.cpp
// (compiled from /root/reference by oracle/ref/Makefile). TEST INFRASTRUCTURE ONLY: pins the initial states and the
// baked -D constants used by the oracle and by the product's generators.
#include "Geometry.hpp"
#include "Utils.hpp"
#include "Target.hpp" // physics/utils/Target.hpp (+ PerlinNoise), unmodified

#include <cstdlib>
#include <cstring>
#include <string>

extern "C" {
long ref_generate_3d_grid(int shape /*0 box, 1 sphere*/, const int res[3], const float start[3], const float end[3], int random,
    float* out_xyzw)
{
  const auto verts = Geometry::Generate3DGrid(shape == 0 ? Geometry::Shape3D::Box : Geometry::Shape3D::Sphere,
      Math::int3(res[0], res[1], res[2]), Math::float3(start[0], start[1], start[2]), Math::float3(end[0], end[1], end[2]),
      random ? Geometry::Distribution::Random : Geometry::Distribution::Uniform);
  for (size_t i = 0; i < verts.size(); ++i)
  {
    out_xyzw[4 * i + 0] = verts[i].x;
    out_xyzw[4 * i + 1] = verts[i].y;
    out_xyzw[4 * i + 2] = verts[i].z;
    out_xyzw[4 * i + 3] = 0.0f;
  }
  return (long)verts.size();
}

long ref_generate_2d_grid(int shape /*0 rectangle, 1 circle*/, int plane /*0 XY, 1 XZ, 2 YZ*/, const int res[2], const float start[3],
    const float end[3], int random, float* out_xyzw)
{
  const Geometry::Plane pl = plane == 0 ? Geometry::Plane::XY : (plane == 1 ? Geometry::Plane::XZ : Geometry::Plane::YZ);
  const auto verts = Geometry::Generate2DGrid(shape == 0 ? Geometry::Shape2D::Rectangle : Geometry::Shape2D::Circle, pl,
      Math::int2(res[0], res[1]), Math::float3(start[0], start[1], start[2]), Math::float3(end[0], end[1], end[2]),
      random ? Geometry::Distribution::Random : Geometry::Distribution::Uniform);
  for (size_t i = 0; i < verts.size(); ++i)
  {
    out_xyzw[4 * i + 0] = verts[i].x;
    out_xyzw[4 * i + 1] = verts[i].y;
    out_xyzw[4 * i + 2] = verts[i].z;
    out_xyzw[4 * i + 3] = 0.0f;
  }
  return (long)verts.size();
}

// the float an OpenCL compiler reads back from "-DNAME=" << Utils::FloatToStr(v)
float ref_baked_constant(float v)
{
  const std::string s = Utils::FloatToStr(v);
  return strtof(s.c_str(), nullptr);
}
void ref_float_to_str(float v, char* out, size_t cap)
{
  const std::string s = Utils::FloatToStr(v);
  strncpy(out, s.c_str(), cap - 1);
  out[cap - 1] = 0;
}
void ref_srand(unsigned seed) { srand(seed); }

// the reference's boids target trajectory (physics/utils/Target.cpp + PerlinNoise.cpp)
void* ref_target_create(unsigned boxSize) { return new Physics::Target(boxSize); }
void ref_target_destroy(void* t) { delete (Physics::Target*)t; }
void ref_target_update(void* t, int dim, float vel, float out[3])
{
  Physics::Target* tg = (Physics::Target*)t;
  tg->updatePos(dim == 3 ? Geometry::Dimension::dim3D : Geometry::Dimension::dim2D, vel);
  const auto p = tg->pos();
  out[0] = p.x;
  out[1] = p.y;
  out[2] = p.z;
}
}
