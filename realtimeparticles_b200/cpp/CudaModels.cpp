// CudaModels.cpp -- host orchestration of the CUDA models: the reference-side binding of librtp_cuda.so. It has to
// reproduce what physics/ocl/{Boids,Fluids,Clouds}.cpp do around their kernels -- JSON keys and parsing rules, the POD
// parameter blocks, the initial states of every preset -- so those values are the reference's (cited); every Context::
// call becomes one C-ABI call, and the presets are data tables (Preset / Lattice / Region below) instead of switch blocks.
#include "CudaModels.hpp"

#include "Geometry.hpp"
#include "Logging.hpp"
#include "Parameters.hpp"
#include "Target.hpp" // physics/utils/Target.hpp (CPU-side Perlin-noise attractor, unchanged)

#include <array>
#include <cstring>
#include <limits>

namespace Physics::CUDA
{
namespace
{
// JSON defaults: same content as initBoidsJson / initFluidsJson / initCloudsJson (Boids.cpp:41-73, Fluids.cpp:52-74,
// Clouds.cpp:67-102), parsed from text so that key order is preserved for the UI
const char* kBoidsJson = R"({"Boids":{"Velocity":[0.5,0.01,5.0],
 "Target":{"Enable##Target":false,"Show":true,"Radius":[2.0,1.0,20.0],"Attract":true},
 "Alignment":{"Enable##Alignment":true,"Scale##Alignment":[1.6,0.0,3.0]},
 "Cohesion":{"Enable##Cohesion":true,"Scale##Cohesion":[1.45,0.0,3.0]},
 "Separation":{"Enable##Separation":true,"Scale##Separation":[1.6,0.0,3.0]}}})";
const char* kFluidsBlock = R"("Fluids":{"Rest Density":[450.0,10.0,1000.0],"Relax CFM":[600.0,100.0,1000.0],
 "Time Step":[0.010,0.0001,0.020],"Nb Jacobi Iterations":[2,1,6],
 "Artificial Pressure":{"Enable##Pressure":true,"Coefficient##Pressure":[0.001,0.0,0.001],"Radius":[0.006,0.001,0.015],"Exp":[4,1,6]},
 "Vorticity Confinement":{"Enable##Vorticity":true,"Coefficient##Vorticity":[0.0004,0.0,0.001],"xSPH Viscosity Coefficient":[0.0001,0.0,0.001]}})";
const char* kCloudsBlock = R"("Clouds":{"Enable Temperature Smoothing":true,"Ground Heat Coefficient":[10.0,0.0,1000.0],
 "Buoyancy Heat Coefficient":[0.10,0.0,5.0],"Gravity Coefficient":[0.0005,0.0,0.1],"Adiabatic Lapse Rate":[5.0,0.0,20.0],
 "Phase Transition Rate":[0.3485,0.0,20.0],"Latent Heat Coefficient":[0.07,0.0,0.100],"Wind Coefficient":[1.0,0.0,1.0]})";

json boidsJson() { return json::parse(kBoidsJson); }
json fluidsJson() { return json::parse(std::string("{") + kFluidsBlock + "}"); }
json cloudsJson() { return json::parse(std::string("{") + kFluidsBlock + "," + kCloudsBlock + "}"); }

void fluidInputsFromJson(const json& fluidsJson, Geometry::Dimension dim, FluidKernelInputs& k)
{
  // Fluids.cpp:225-243 / Clouds.cpp:291-305
  k.restDensity = (float)(fluidsJson["Rest Density"][0]);
  k.relaxCFM = (float)(fluidsJson.at("Relax CFM")[0]);
  k.timeStep = (float)(fluidsJson["Time Step"][0]);
  k.dim = (uint32_t)((dim == Geometry::Dimension::dim2D) ? 2 : 3);
  k.isArtPressureEnabled = (uint32_t)((fluidsJson["Artificial Pressure"]["Enable##Pressure"] == true) ? 1 : 0);
  k.artPressureCoeff = (float)(fluidsJson["Artificial Pressure"]["Coefficient##Pressure"][0]);
  k.artPressureRadius = (float)(fluidsJson["Artificial Pressure"]["Radius"][0]);
  k.artPressureExp = (uint32_t)(fluidsJson["Artificial Pressure"]["Exp"][0]);
  k.isVorticityConfEnabled = (uint32_t)((fluidsJson["Vorticity Confinement"]["Enable##Vorticity"] == true) ? 1 : 0);
  k.vorticityConfCoeff = (float)(fluidsJson["Vorticity Confinement"]["Coefficient##Vorticity"][0]);
  k.xsphViscosityCoeff = (float)(fluidsJson["Vorticity Confinement"]["xSPH Viscosity Coefficient"][0]);
}
} // namespace

template <typename... K>
void CudaModel<K...>::uploadParticles(const std::vector<Math::float3>& verts, bool velocityIsPosition, const float* colour)
{
  // Fluids.cpp:383-395 / Boids.cpp:310-318: +inf tail, blocking uploads
  const float inf = std::numeric_limits<float>::infinity();
  std::vector<std::array<float, 4>> pos(m_maxNbParticles, std::array<float, 4>({ inf, inf, inf, 0.0f }));
  for (size_t i = 0; i < verts.size() && i < pos.size(); ++i)
    pos[i] = { verts[i].x, verts[i].y, verts[i].z, 0.0f };
  rtp_upload(m_handle, RTP_F_POS, pos.data(), 16 * pos.size());
  if (velocityIsPosition)
  {
    rtp_upload(m_handle, RTP_F_VEL, pos.data(), 16 * pos.size());
  }
  else
  {
    std::vector<std::array<float, 4>> vel(m_maxNbParticles, std::array<float, 4>({ 0.0f, 0.0f, 0.0f, 0.0f }));
    rtp_upload(m_handle, RTP_F_VEL, vel.data(), 16 * vel.size());
  }
  if (colour)
  {
    std::vector<std::array<float, 4>> col(m_maxNbParticles, std::array<float, 4>({ colour[0], colour[1], colour[2], colour[3] }));
    rtp_upload(m_handle, RTP_F_COL, col.data(), 16 * col.size());
  }
}

// ------------------------------------------------------------------ Boids (physics/ocl/Boids.cpp)

Boids::Boids(ModelParams params)
    : CudaModel<BoidsRuleKernelInputs, TargetKernelInputs>(params, RTP_MODEL_BOIDS, 3000 /* Boids.cpp:78 */,
        BoidsRuleKernelInputs { 0.5f, 1.6f, 1.6f, 1.45f }, TargetKernelInputs { 2.0f, 1 }, boidsJson())
    , m_target(std::make_unique<Target>(params.boxSize.x))
{
  m_init = m_handle != nullptr;
  reset();
}
Boids::~Boids() = default;
Math::float3 Boids::targetPos() const { return m_target->pos(); }
bool Boids::isTargetActivated() const { return m_target->isActivated(); }
bool Boids::isTargetVisible() const { return m_target->isVisible(); }

void Boids::transferJsonInputsToModel(json& inputJson)
{
  if (!m_init)
    return;
  try
  { // Boids.cpp:182-203
    const auto& boidsJson = inputJson["Boids"];
    auto& rules = getKernelInput<BoidsRuleKernelInputs>(0);
    rules.velocityScale = (float)(boidsJson["Velocity"][0]);
    rules.alignmentScale = boidsJson["Alignment"]["Enable##Alignment"] ? (float)(boidsJson["Alignment"]["Scale##Alignment"][0]) : 0.0f;
    rules.separationScale = boidsJson["Separation"]["Enable##Separation"] ? (float)(boidsJson["Separation"]["Scale##Separation"][0]) : 0.0f;
    rules.cohesionScale = boidsJson["Cohesion"]["Enable##Cohesion"] ? (float)(boidsJson["Cohesion"]["Scale##Cohesion"][0]) : 0.0f;
    auto& target = getKernelInput<TargetKernelInputs>(1);
    m_target->activate(boidsJson["Target"]["Enable##Target"]);
    m_target->show(boidsJson["Target"]["Show"]);
    m_target->setRadiusEffect(boidsJson["Target"]["Radius"][0]);
    target.targetRadiusEffect = m_target->radiusEffect();
    m_target->setSignEffect((int)(boidsJson["Target"]["Attract"]));
    target.targetSignEffect = m_target->signEffect();
  }
  catch (...)
  {
    LOG_ERROR("Boids Input Json parsing is incorrect, did you use a wrong path for a parameter?");
    throw std::runtime_error("Wrong Json parsing");
  }
}

void Boids::transferKernelInputsToGPU()
{
  if (!m_init)
    return;
  const auto t = m_target->pos();
  const float tp[4] = { t.x, t.y, t.z, 0.0f };
  rtp_set_boids_params(m_handle, &getKernelInput<BoidsRuleKernelInputs>(0), &getKernelInput<TargetKernelInputs>(1), tp,
      isTargetActivated() ? 1 : 0);
}

// ---- initial states as data: a lattice is a particle count, a shape and a region given per axis as (num * box) / den --
// the reference spells its regions as box fractions (Boids.cpp:277-321, Fluids.cpp:273-398, Clouds.cpp:401-501); keeping the
// operation (one multiplication by a small integer, one division) keeps the corner coordinates bit-identical.
namespace
{
struct Frac
{
  float num, den;
  float of(float box) const { return num * box / den; }
};
constexpr Frac ZERO { 0.0f, 1.0f };
struct Region
{
  Frac lo[3], hi[3];
  Math::float3 start(const Geometry::BoxSize3D& b) const { return { lo[0].of((float)b.x), lo[1].of((float)b.y), lo[2].of((float)b.z) }; }
  Math::float3 end(const Geometry::BoxSize3D& b) const { return { hi[0].of((float)b.x), hi[1].of((float)b.y), hi[2].of((float)b.z) }; }
};
enum LatticeShape
{
  LATTICE_BOX, // 3D box / 2D rectangle
  LATTICE_ROUND // 3D sphere / 2D circle
};
struct Lattice
{
  int count; // Utils::NbParticles, 0 = no such lattice; the subdivision follows from it unless res is given
  LatticeShape shape;
  Region region;
  int res[3]; // explicit subdivision (2D: res[0], res[1]) or { 0, 0, 0 }
};
struct Preset
{
  Utils::PhysicsCase pcase;
  Lattice main, extra; // extra: a second lattice appended to the first (the pool under the drop)
};

std::vector<Math::float3> generate(const Lattice& l, bool dim2D, const Geometry::BoxSize3D& box,
    Geometry::Distribution dist = Geometry::Distribution::Uniform)
{
  const Math::float3 a = l.region.start(box), b = l.region.end(box);
  if (dim2D)
  {
    Math::int2 res = { l.res[0], l.res[1] };
    if (!l.res[0])
    {
      const auto& sub = Utils::GetNbParticlesSubdiv2D((Utils::NbParticles)l.count);
      res = { sub[0], sub[1] };
    }
    return Geometry::Generate2DGrid(l.shape == LATTICE_ROUND ? Geometry::Shape2D::Circle : Geometry::Shape2D::Rectangle, Geometry::Plane::YZ, res, a, b, dist);
  }
  Math::int3 res = { l.res[0], l.res[1], l.res[2] };
  if (!l.res[0])
  {
    const auto& sub = Utils::GetNbParticlesSubdiv3D((Utils::NbParticles)l.count);
    res = { sub[0], sub[1], sub[2] };
  }
  return Geometry::Generate3DGrid(l.shape == LATTICE_ROUND ? Geometry::Shape3D::Sphere : Geometry::Shape3D::Box, res, a, b, dist);
}

// particles of a preset (main lattice + optional extra one); nb = their nominal number
std::vector<Math::float3> generate(const Preset& p, bool dim2D, const Geometry::BoxSize3D& box, size_t& nb,
    Geometry::Distribution dist = Geometry::Distribution::Uniform)
{
  std::vector<Math::float3> verts = generate(p.main, dim2D, box, dist);
  nb = (size_t)p.main.count;
  if (p.extra.count)
  {
    const auto more = generate(p.extra, dim2D, box, dist);
    verts.insert(verts.end(), more.begin(), more.end());
    nb += (size_t)p.extra.count;
  }
  return verts;
}

template <size_t N>
const Preset* findPreset(const Preset (&table)[N], Utils::PhysicsCase pcase)
{
  for (const Preset& p : table)
    if (p.pcase == pcase)
      return &p;
  return nullptr;
}

constexpr Lattice NONE { 0, LATTICE_BOX, { { ZERO, ZERO, ZERO }, { ZERO, ZERO, ZERO } }, { 0, 0, 0 } };
using NP = Utils::NbParticles;
using PC = Utils::PhysicsCase;

// boids: a ball (3D) / disc (2D, YZ plane) of a third of the box around the centre; the case only sets the count
constexpr Region BOIDS_3D { { { 1, -6 }, { 1, -6 }, { 1, -6 } }, { { 1, 6 }, { 1, 6 }, { 1, 6 } } };
constexpr Region BOIDS_2D { { ZERO, { 1, -6 }, { 1, -6 } }, { ZERO, { 1, 6 }, { 1, 6 } } };
constexpr struct { PC pcase; int count; } BOIDS_COUNTS[] = { { PC::BOIDS_SMALL, NP::P512 }, { PC::BOIDS_MEDIUM, NP::P16K },
  { PC::BOIDS_LARGE, NP::P65K }, { PC::BOIDS_XLARGE, NP::P130K } };

// fluids: dam (the -x half of the lower half), bomb (a ball in the centre), drop (a block above a pool)
constexpr Preset FLUIDS_3D[] = {
  { PC::FLUIDS_DAM, { NP::P130K, LATTICE_BOX, { { { 1, -2 }, { 1, -2 }, { 1, -2 } }, { { 1, 2 }, ZERO, ZERO } }, { 0, 0, 0 } }, NONE },
  { PC::FLUIDS_BOMB, { NP::P65K, LATTICE_ROUND, BOIDS_3D, { 0, 0, 0 } }, NONE },
  { PC::FLUIDS_DROP, { NP::P4K, LATTICE_BOX, { { { 1, -10 }, { 2, 10 }, { 1, -10 } }, { { 1, 10 }, { 4, 10 }, { 1, 10 } } }, { 0, 0, 0 } },
      { NP::P65K, LATTICE_BOX, { { { 1, -2 }, { 1, -2 }, { 1, -2 } }, { { 1, 2 }, { 1, -2.55f }, { 1, 2 } } }, { 64, 16, 64 } } },
};
constexpr Preset FLUIDS_2D[] = {
  { PC::FLUIDS_DAM, { NP::P4K, LATTICE_BOX, { { ZERO, { 1, -2 }, { 1, -2 } }, { ZERO, ZERO, ZERO } }, { 0, 0, 0 } }, NONE },
  { PC::FLUIDS_BOMB, { NP::P4K, LATTICE_BOX, BOIDS_2D, { 0, 0, 0 } }, NONE },
  { PC::FLUIDS_DROP, { NP::P512, LATTICE_BOX, { { ZERO, { 2, 10 }, { 1, -10 } }, { ZERO, { 4, 10 }, { 1, 10 } } }, { 0, 0, 0 } },
      { NP::P4K, LATTICE_BOX, { { ZERO, { 1, -2 }, { 1, -2 } }, { ZERO, ZERO, { 1, 2 } } }, { 64, 128, 0 } } },
};

// clouds: randomly filled slab at the bottom (cumulus) or the whole box (homogeneous)
constexpr Preset CLOUDS_3D[] = {
  { PC::CLOUDS_CUMULUS, { NP::P65K, LATTICE_BOX, { { { 1, -2 }, { 1, -2 }, { 1, -2 } }, { { 1, 2 }, { 1, -4 }, { 1, 2 } } }, { 0, 0, 0 } }, NONE },
  { PC::CLOUDS_HOMOGENEOUS, { NP::P65K, LATTICE_BOX, { { { 1, -2 }, { 1, -2 }, { 1, -2 } }, { { 1, 2 }, { 1, 2 }, { 1, 2 } } }, { 0, 0, 0 } }, NONE },
};
constexpr Preset CLOUDS_2D[] = {
  { PC::CLOUDS_CUMULUS, { NP::P8K, LATTICE_BOX, { { ZERO, { 1, -2 }, { 1, -2 } }, { ZERO, ZERO, { 1, 2 } } }, { 0, 0, 0 } }, NONE },
  { PC::CLOUDS_HOMOGENEOUS, { NP::P8K, LATTICE_BOX, { { ZERO, { 1, -2 }, { 1, -2 } }, { ZERO, { 1, 2 }, { 1, 2 } } }, { 0, 0, 0 } }, NONE },
};
} // namespace

void Boids::reset()
{
  if (!m_init)
    return;
  resetInputJson(boidsJson());
  for (const auto& c : BOIDS_COUNTS) // the case only chooses the flock's size (Boids.cpp:235-257)
    if (c.pcase == m_case)
      m_currNbParticles = (size_t)c.count;
  rtp_set_nb_particles(m_handle, m_currNbParticles);
  rtp_set_dimension(m_handle, m_dimension == Geometry::Dimension::dim2D ? 2 : 3);
  json js = getInputJson();
  updateModelWithInputJson(js);
  initBoidsParticles();
  rtp_reset_ids(m_handle);
}

void Boids::initBoidsParticles()
{ // Boids.cpp:277-321
  if (m_currNbParticles > m_maxNbParticles)
  {
    LOG_ERROR("Cannot init boids, current number of particles is higher than max limit");
    return;
  }
  const bool dim2D = m_dimension == Geometry::Dimension::dim2D;
  const Lattice flock { (int)m_currNbParticles, LATTICE_ROUND, dim2D ? BOIDS_2D : BOIDS_3D, { 0, 0, 0 } };
  const std::vector<Math::float3> gridVerts = generate(flock, dim2D, m_boxSize);
  const float colour[4] = { 1.0f, 0.02f, 0.02f, 0.5f }; // bd_fillBoidsColor boids.cl:38-41
  uploadParticles(gridVerts, true /* "Using same buffer to initialize vel", Boids.cpp:316-318 */, colour);
}

void Boids::update()
{ // Boids.cpp:323-384
  if (!m_init)
    return;
  rtp_set_boundary(m_handle, m_boundary == Boundary::CyclicWall ? RTP_BOUNDARY_CYCLIC_WALL : RTP_BOUNDARY_BOUNCING_WALL);
  const bool movingTarget = !m_pause && isTargetActivated();
  if (movingTarget)
  {
    m_target->updatePos(m_dimension, getKernelInput<BoidsRuleKernelInputs>(0).velocityScale);
    transferKernelInputsToGPU();
  }
  stepDevice(!movingTarget);
}

// ------------------------------------------------------------------ Fluids (physics/ocl/Fluids.cpp)

Fluids::Fluids(ModelParams params)
    : CudaModel<FluidKernelInputs>(params, RTP_MODEL_FLUIDS, 100 /* Fluids.cpp:79 */,
        FluidKernelInputs { 450.0f, 600.0f, 0.010f, 3, 1, 0.006f, 0.001f, 4, 1, 0.0004f, 0.0001f }, fluidsJson())
{
  m_init = m_handle != nullptr;
  reset();
}

void Fluids::transferJsonInputsToModel(json& inputJson)
{
  if (!m_init || inputJson.empty())
    return;
  try
  {
    const auto& fluidsJson = inputJson["Fluids"];
    m_nbJacobiIters = fluidsJson["Nb Jacobi Iterations"][0];
    fluidInputsFromJson(fluidsJson, m_dimension, getKernelInput<FluidKernelInputs>(0));
  }
  catch (...)
  {
    LOG_ERROR("Fluids Input Json parsing is incorrect, did you use a wrong path for a parameter?");
    throw std::runtime_error("Wrong Json parsing");
  }
}

void Fluids::transferKernelInputsToGPU()
{
  if (!m_init)
    return;
  rtp_set_fluid_params(m_handle, &getKernelInput<FluidKernelInputs>(0), (int)m_nbJacobiIters);
}

void Fluids::reset()
{ // Fluids.cpp:196-216
  if (!m_init)
    return;
  resetInputJson(fluidsJson());
  rtp_set_dimension(m_handle, m_dimension == Geometry::Dimension::dim2D ? 2 : 3);
  json js = getInputJson();
  updateModelWithInputJson(js);
  initFluidsParticles();
  rtp_reset_ids(m_handle);
}

void Fluids::initFluidsParticles()
{ // Fluids.cpp:273-398
  const bool dim2D = m_dimension == Geometry::Dimension::dim2D;
  const Preset* preset = dim2D ? findPreset(FLUIDS_2D, m_case) : findPreset(FLUIDS_3D, m_case);
  std::vector<Math::float3> gridVerts;
  if (!preset)
  {
    LOG_ERROR("Unkown case type");
    // (the reference falls through with an empty region: every particle of the current count at the origin)
    const Lattice origin { (int)m_currNbParticles, LATTICE_BOX, NONE.region, { 0, 0, 0 } };
    gridVerts = generate(origin, dim2D, m_boxSize);
  }
  else
    gridVerts = generate(*preset, dim2D, m_boxSize, m_currNbParticles);
  if (m_currNbParticles > m_maxNbParticles)
    m_currNbParticles = m_maxNbParticles;
  rtp_set_nb_particles(m_handle, m_currNbParticles);
  const float colour[4] = { 0.0f, 0.1f, 1.0f, 0.0f };
  uploadParticles(gridVerts, false, colour);
}

void Fluids::update()
{ // Fluids.cpp:400-471
  if (!m_init)
    return;
  stepDevice();
}

// ------------------------------------------------------------------ Clouds (physics/ocl/Clouds.cpp)

Clouds::Clouds(ModelParams params)
    : CudaModel<FluidKernelInputs, CloudKernelInputs>(params, RTP_MODEL_CLOUDS, 100 /* Clouds.cpp:107 */,
        FluidKernelInputs { 450.0f, 600.0f, 0.010f, 3, 1, 0.006f, 0.001f, 4, 1, 0.0004f, 0.0001f },
        CloudKernelInputs { 3, 0.01f, 400.0f, 10.0f, 0.10f, 0.0005f, 5.0f, 0.3485f, 0.07f, 1, 600.0f, 0.75f, 1.0f }, cloudsJson())
{
  // Clouds.cpp:202-213
  auto add = [&](const char* name, const char* buffer, std::pair<float, float> s, std::pair<float, float> u)
  { m_allDisplayableQuantities.insert(std::make_pair(std::string(name), PhysicalQuantity { name, buffer, s, u })); };
  add("Particle ID", "p_partID", { 0.0f, (float)(m_maxNbParticles - 1) }, { 0.0f, (float)(32000 - 1) });
  add("Vapor Density", "p_vaporDens", { 0.0f, 100.0f }, { 0.001f, 100.0f });
  add("Cloud Density", "p_cloudDens", { 0.0f, 100.0f }, { 1.0f, 15.0f });
  add("Net Force", "p_buoyancy", { -10.0f, 10.0f }, { -1.0f, 1.0f });
  add("Temperature", "p_temp", { 0.0f, 500.0f }, { 223.0f, 293.0f });
  m_currentDisplayedQuantityName = "Cloud Density";
  m_init = m_handle != nullptr && !m_allDisplayableQuantities.empty();
  reset();
}

void Clouds::transferJsonInputsToModel(json& inputJson)
{
  if (!m_init)
    return;
  try
  { // Clouds.cpp:286-322
    const auto& fluidsJson = inputJson["Fluids"];
    m_nbJacobiIters = fluidsJson["Nb Jacobi Iterations"][0];
    fluidInputsFromJson(fluidsJson, m_dimension, getKernelInput<FluidKernelInputs>(0));
    const auto& cloudsJson = inputJson["Clouds"];
    auto& c = getKernelInput<CloudKernelInputs>(1);
    c.restDensity = (float)(fluidsJson["Rest Density"][0]);
    c.timeStep = (float)(fluidsJson["Time Step"][0]);
    c.dim = (uint32_t)((m_dimension == Geometry::Dimension::dim2D) ? 2 : 3);
    c.relaxCFM = (float)(fluidsJson["Relax CFM"][0]);
    c.isTempSmoothingEnabled = (uint32_t)(cloudsJson["Enable Temperature Smoothing"] ? 1 : 0);
    c.groundHeatCoeff = (float)(cloudsJson["Ground Heat Coefficient"][0]);
    c.buoyancyCoeff = (float)(cloudsJson["Buoyancy Heat Coefficient"][0]);
    c.gravCoeff = (float)(cloudsJson["Gravity Coefficient"][0]);
    c.adiabaticLapseRate = (float)(cloudsJson["Adiabatic Lapse Rate"][0]);
    c.phaseTransitionRate = (float)(cloudsJson["Phase Transition Rate"][0]);
    c.latentHeatCoeff = (float)(cloudsJson["Latent Heat Coefficient"][0]);
    c.windCoeff = (float)(cloudsJson["Wind Coefficient"][0]);
  }
  catch (...)
  {
    LOG_ERROR("Clouds Input Json parsing is incorrect, did you use a wrong path for a parameter?");
    throw std::runtime_error("Wrong Json parsing");
  }
}

void Clouds::transferKernelInputsToGPU()
{
  if (!m_init)
    return;
  rtp_set_fluid_params(m_handle, &getKernelInput<FluidKernelInputs>(0), (int)m_nbJacobiIters);
  rtp_set_cloud_params(m_handle, &getKernelInput<CloudKernelInputs>(1));
}

void Clouds::pushDisplayedQuantity()
{ // Clouds.cpp:610-616
  static const std::map<std::string, int> fields = { { "p_partID", RTP_F_PART_ID }, { "p_vaporDens", RTP_F_VAPOR_DENS },
    { "p_cloudDens", RTP_F_CLOUD_DENS }, { "p_buoyancy", RTP_F_BUOYANCY }, { "p_temp", RTP_F_TEMP } };
  const auto& q = currentDisplayedPhysicalQuantity();
  const auto it = fields.find(q.bufferName);
  if (it != fields.end())
    rtp_set_displayed_quantity(m_handle, it->second, q.userRange.first, q.userRange.second);
}

void Clouds::reset()
{ // Clouds.cpp:384-399
  if (!m_init)
    return;
  resetInputJson(cloudsJson());
  rtp_set_dimension(m_handle, m_dimension == Geometry::Dimension::dim2D ? 2 : 3);
  json js = getInputJson();
  updateModelWithInputJson(js);
  initCloudsParticles();
  rtp_reset_ids(m_handle);
}

void Clouds::initCloudsParticles()
{ // Clouds.cpp:401-501
  const bool dim2D = m_dimension == Geometry::Dimension::dim2D;
  const Preset* preset = dim2D ? findPreset(CLOUDS_2D, m_case) : findPreset(CLOUDS_3D, m_case);
  if (!preset)
  {
    LOG_ERROR("Unkown case type");
    preset = dim2D ? &CLOUDS_2D[1] : &CLOUDS_3D[1]; // (the reference treats every other case like the homogeneous one)
  }
  const std::vector<Math::float3> gridVerts = generate(*preset, dim2D, m_boxSize, m_currNbParticles, Geometry::Distribution::Random);
  if (m_currNbParticles > m_maxNbParticles)
    m_currNbParticles = m_maxNbParticles;
  rtp_set_nb_particles(m_handle, m_currNbParticles);
  const float colour[4] = { 0.0f, 0.1f, 1.0f, 0.0f };
  uploadParticles(gridVerts, false, colour);
  std::vector<float> cloudDens(m_maxNbParticles, 0.0f);
  rtp_upload(m_handle, RTP_F_CLOUD_DENS, cloudDens.data(), 4 * cloudDens.size());
  std::vector<float> partID(m_maxNbParticles, 0.0f);
  for (size_t i = 0; i != partID.size(); ++i)
    partID[i] = (float)i;
  rtp_upload(m_handle, RTP_F_PART_ID, partID.data(), 4 * partID.size());
  rtp_init_clouds_fields(m_handle); // cld_initTemperature + cld_initVaporDensity, Clouds.cpp:495-497
}

void Clouds::update()
{ // Clouds.cpp:503-627
  if (!m_init)
    return;
  pushDisplayedQuantity();
  stepDevice();
}

std::unique_ptr<Model> CreateModel(ModelType type, ModelParams params)
{ // physics/Model.cpp:10-24 with the CUDA classes
  switch ((int)type)
  {
  case ModelType::BOIDS: return std::make_unique<Boids>(params);
  case ModelType::FLUIDS: return std::make_unique<Fluids>(params);
  case ModelType::CLOUDS: return std::make_unique<Clouds>(params);
  default: return nullptr;
  }
}
} // namespace Physics::CUDA

// ------------------------------------------------------------------ flat C entry points for the headless harness / tests
extern "C" {
__attribute__((visibility("default"))) void* rtpm_create(int type, uint64_t maxParticles, int pCase, int dim3, const uint32_t box[3],
    const uint32_t grid[3])
{
  Physics::ModelParams p;
  p.maxNbParticles = maxParticles;
  p.currNbParticles = maxParticles;
  p.boxSize = { box[0], box[1], box[2] };
  p.gridRes = { grid[0], grid[1], grid[2] };
  p.dimension = dim3 ? Geometry::Dimension::dim3D : Geometry::Dimension::dim2D;
  p.pCase = (Utils::PhysicsCase)pCase;
  auto m = Physics::CUDA::CreateModel((Physics::ModelType)type, p);
  return m.release();
}
__attribute__((visibility("default"))) void rtpm_destroy(void* m) { delete (Physics::Model*)m; }
__attribute__((visibility("default"))) int rtpm_is_init(void* m) { return ((Physics::Model*)m)->isInit() ? 1 : 0; }
__attribute__((visibility("default"))) uint64_t rtpm_nb_particles(void* m) { return ((Physics::Model*)m)->nbParticles(); }
__attribute__((visibility("default"))) void rtpm_update(void* m) { ((Physics::Model*)m)->update(); }
__attribute__((visibility("default"))) void rtpm_reset(void* m) { ((Physics::Model*)m)->reset(); }
__attribute__((visibility("default"))) void rtpm_pause(void* m, int p) { ((Physics::Model*)m)->pause(p != 0); }
__attribute__((visibility("default"))) void rtpm_set_boundary(void* m, int cyclic)
{
  ((Physics::Model*)m)->setBoundary(cyclic ? Physics::Boundary::CyclicWall : Physics::Boundary::BouncingWall);
}
// merge-patch the model's JSON blob with a JSON text (what ui/PhysicsWidget does every frame); returns 0, or -1 when the
// model threw "Wrong Json parsing"
__attribute__((visibility("default"))) int rtpm_update_input_json(void* m, const char* text)
{
  try
  {
    json cur = ((Physics::Model*)m)->getInputJson();
    cur.merge_patch(json::parse(text));
    ((Physics::Model*)m)->updateInputJson(cur);
    return 0;
  }
  catch (...)
  {
    return -1;
  }
}
__attribute__((visibility("default"))) int rtpm_get_input_json(void* m, char* out, size_t cap)
{
  const std::string s = ((Physics::Model*)m)->getInputJson().dump();
  if (s.size() + 1 > cap)
    return -(int)s.size();
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}
// the rtp_handle behind a model, for rtp_download / rtp_sync on its buffers
__attribute__((visibility("default"))) void* rtpm_handle(void* m, int type)
{
  switch (type)
  {
  case 0: return ((Physics::CUDA::Boids*)(Physics::Model*)m)->handle();
  case 1: return ((Physics::CUDA::Fluids*)(Physics::Model*)m)->handle();
  case 2: return ((Physics::CUDA::Clouds*)(Physics::Model*)m)->handle();
  }
  return nullptr;
}
__attribute__((visibility("default"))) void rtpm_set_step_flags(void* m, int type, unsigned flags)
{
  switch (type)
  {
  case 0: ((Physics::CUDA::Boids*)(Physics::Model*)m)->setStepFlags(flags); break;
  case 1: ((Physics::CUDA::Fluids*)(Physics::Model*)m)->setStepFlags(flags); break;
  case 2: ((Physics::CUDA::Clouds*)(Physics::Model*)m)->setStepFlags(flags); break;
  }
}
}
