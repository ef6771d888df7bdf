// CudaModels.cpp -- host orchestration of the CUDA models; mirrors physics/ocl/{Boids,Fluids,Clouds}.cpp with every
// Context:: call replaced by one C-ABI call. Presets and JSON handling follow the reference line by line (cited).
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

void Boids::reset()
{
  if (!m_init)
    return;
  resetInputJson(boidsJson());
  switch (m_case)
  { // Boids.cpp:235-257
  case Utils::PhysicsCase::BOIDS_SMALL: m_currNbParticles = Utils::NbParticles::P512; break;
  case Utils::PhysicsCase::BOIDS_MEDIUM: m_currNbParticles = Utils::NbParticles::P16K; break;
  case Utils::PhysicsCase::BOIDS_LARGE: m_currNbParticles = Utils::NbParticles::P65K; break;
  case Utils::PhysicsCase::BOIDS_XLARGE: m_currNbParticles = Utils::NbParticles::P130K; break;
  default: break;
  }
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
  std::vector<Math::float3> gridVerts;
  if (m_dimension == Geometry::Dimension::dim2D)
  {
    const auto& subdiv2D = Utils::GetNbParticlesSubdiv2D((Utils::NbParticles)m_currNbParticles);
    Math::int2 grid2DRes = { subdiv2D[0], subdiv2D[1] };
    Math::float3 start2D = { 0.0f, m_boxSize.y / -6.0f, m_boxSize.z / -6.0f };
    Math::float3 end2D = { 0.0f, m_boxSize.y / 6.0f, m_boxSize.z / 6.0f };
    gridVerts = Geometry::Generate2DGrid(Geometry::Shape2D::Circle, Geometry::Plane::YZ, grid2DRes, start2D, end2D);
  }
  else
  {
    const auto& subdiv3D = Utils::GetNbParticlesSubdiv3D((Utils::NbParticles)m_currNbParticles);
    Math::int3 grid3DRes = { subdiv3D[0], subdiv3D[1], subdiv3D[2] };
    Math::float3 start3D = { m_boxSize.x / -6.0f, m_boxSize.y / -6.0f, m_boxSize.z / -6.0f };
    Math::float3 end3D = { m_boxSize.x / 6.0f, m_boxSize.y / 6.0f, m_boxSize.z / 6.0f };
    gridVerts = Geometry::Generate3DGrid(Geometry::Shape3D::Sphere, grid3DRes, start3D, end3D);
  }
  const float colour[4] = { 1.0f, 0.02f, 0.02f, 0.5f }; // bd_fillBoidsColor boids.cl:38-41
  uploadParticles(gridVerts, true /* "Using same buffer to initialize vel", Boids.cpp:316-318 */, colour);
}

void Boids::update()
{ // Boids.cpp:323-384
  if (!m_init)
    return;
  rtp_set_boundary(m_handle, m_boundary == Boundary::CyclicWall ? RTP_BOUNDARY_CYCLIC_WALL : RTP_BOUNDARY_BOUNCING_WALL);
  if (!m_pause && isTargetActivated())
  {
    m_target->updatePos(m_dimension, getKernelInput<BoidsRuleKernelInputs>(0).velocityScale);
    transferKernelInputsToGPU();
  }
  stepDevice();
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
  std::vector<Math::float3> gridVerts;
  Math::float3 startFluidPos = { 0.0f, 0.0f, 0.0f };
  Math::float3 endFluidPos = { 0.0f, 0.0f, 0.0f };
  if (m_dimension == Geometry::Dimension::dim2D)
  {
    switch (m_case)
    {
    case Utils::PhysicsCase::FLUIDS_DAM:
      m_currNbParticles = Utils::NbParticles::P4K;
      startFluidPos = { 0.0f, m_boxSize.y / -2.0f, m_boxSize.z / -2.0f };
      endFluidPos = { 0.0f, 0.0f, 0.0f };
      break;
    case Utils::PhysicsCase::FLUIDS_BOMB:
      m_currNbParticles = Utils::NbParticles::P4K;
      startFluidPos = { 0.0f, m_boxSize.y / -6.0f, m_boxSize.z / -6.0f };
      endFluidPos = { 0.0f, m_boxSize.y / 6.0f, m_boxSize.z / 6.0f };
      break;
    case Utils::PhysicsCase::FLUIDS_DROP:
      m_currNbParticles = Utils::NbParticles::P512;
      startFluidPos = { 0.0f, 2.0f * m_boxSize.y / 10.0f, m_boxSize.z / -10.0f };
      endFluidPos = { 0.0f, 4.0f * m_boxSize.y / 10.0f, m_boxSize.z / 10.0f };
      break;
    default: LOG_ERROR("Unkown case type"); break;
    }
    const auto& subdiv2D = Utils::GetNbParticlesSubdiv2D((Utils::NbParticles)m_currNbParticles);
    Math::int2 grid2DRes = { subdiv2D[0], subdiv2D[1] };
    gridVerts = Geometry::Generate2DGrid(Geometry::Shape2D::Rectangle, Geometry::Plane::YZ, grid2DRes, startFluidPos, endFluidPos);
    if (m_case == Utils::PhysicsCase::FLUIDS_DROP)
    {
      m_currNbParticles += Utils::NbParticles::P4K;
      Math::int2 res2 = { 64, 128 };
      startFluidPos = { 0.0f, m_boxSize.y / -2.0f, m_boxSize.z / -2.0f };
      endFluidPos = { 0.0f, 0.0f, m_boxSize.z / 2.0f };
      auto bottom = Geometry::Generate2DGrid(Geometry::Shape2D::Rectangle, Geometry::Plane::YZ, res2, startFluidPos, endFluidPos);
      gridVerts.insert(gridVerts.end(), bottom.begin(), bottom.end());
    }
  }
  else
  {
    Geometry::Shape3D shape = Geometry::Shape3D::Box;
    switch (m_case)
    {
    case Utils::PhysicsCase::FLUIDS_DAM:
      m_currNbParticles = Utils::NbParticles::P130K;
      startFluidPos = { m_boxSize.x / -2.0f, m_boxSize.y / -2.0f, m_boxSize.z / -2.0f };
      endFluidPos = { m_boxSize.x / 2.0f, 0.0f, 0.0f };
      break;
    case Utils::PhysicsCase::FLUIDS_BOMB:
      m_currNbParticles = Utils::NbParticles::P65K;
      shape = Geometry::Shape3D::Sphere;
      startFluidPos = { m_boxSize.x / -6.0f, m_boxSize.y / -6.0f, m_boxSize.z / -6.0f };
      endFluidPos = { m_boxSize.x / 6.0f, m_boxSize.y / 6.0f, m_boxSize.z / 6.0f };
      break;
    case Utils::PhysicsCase::FLUIDS_DROP:
      m_currNbParticles = Utils::NbParticles::P4K;
      startFluidPos = { m_boxSize.x / -10.0f, 2.0f * m_boxSize.y / 10.0f, m_boxSize.z / -10.0f };
      endFluidPos = { m_boxSize.x / 10.0f, 4.0f * m_boxSize.y / 10.0f, m_boxSize.z / 10.0f };
      break;
    default: LOG_ERROR("Unkown case type"); break;
    }
    const auto& subdiv3D = Utils::GetNbParticlesSubdiv3D((Utils::NbParticles)m_currNbParticles);
    Math::int3 grid3DRes = { subdiv3D[0], subdiv3D[1], subdiv3D[2] };
    gridVerts = Geometry::Generate3DGrid(shape, grid3DRes, startFluidPos, endFluidPos);
    if (m_case == Utils::PhysicsCase::FLUIDS_DROP)
    {
      m_currNbParticles += Utils::NbParticles::P65K;
      Math::int3 res3 = { 64, 16, 64 };
      startFluidPos = { m_boxSize.x / -2.0f, m_boxSize.y / -2.0f, m_boxSize.z / -2.0f };
      endFluidPos = { m_boxSize.x / 2.0f, m_boxSize.y / -2.55f, m_boxSize.z / 2.0f };
      auto bottom = Geometry::Generate3DGrid(Geometry::Shape3D::Box, res3, startFluidPos, endFluidPos);
      gridVerts.insert(gridVerts.end(), bottom.begin(), bottom.end());
    }
  }
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
  std::vector<Math::float3> gridVerts;
  Math::float3 startFluidPos = { 0.0f, 0.0f, 0.0f };
  Math::float3 endFluidPos = { 0.0f, 0.0f, 0.0f };
  const Geometry::Distribution distribution = Geometry::Distribution::Random;
  const bool cumulus = m_case == Utils::PhysicsCase::CLOUDS_CUMULUS;
  if (m_case != Utils::PhysicsCase::CLOUDS_CUMULUS && m_case != Utils::PhysicsCase::CLOUDS_HOMOGENEOUS)
    LOG_ERROR("Unkown case type");
  if (m_dimension == Geometry::Dimension::dim2D)
  {
    m_currNbParticles = Utils::NbParticles::P8K;
    startFluidPos = { 0.0f, m_boxSize.y / -2.0f, m_boxSize.z / -2.0f };
    endFluidPos = { 0.0f, cumulus ? 0.0f : m_boxSize.y / 2.0f, m_boxSize.z / 2.0f };
    const auto& subdiv2D = Utils::GetNbParticlesSubdiv2D((Utils::NbParticles)m_currNbParticles);
    Math::int2 grid2DRes = { subdiv2D[0], subdiv2D[1] };
    gridVerts = Geometry::Generate2DGrid(Geometry::Shape2D::Rectangle, Geometry::Plane::YZ, grid2DRes, startFluidPos, endFluidPos, distribution);
  }
  else
  {
    m_currNbParticles = Utils::NbParticles::P65K;
    startFluidPos = { m_boxSize.x / -2.0f, m_boxSize.y / -2.0f, m_boxSize.z / -2.0f };
    endFluidPos = { m_boxSize.x / 2.0f, cumulus ? m_boxSize.y / -4.0f : m_boxSize.y / 2.0f, m_boxSize.z / 2.0f };
    const auto& subdiv3D = Utils::GetNbParticlesSubdiv3D((Utils::NbParticles)m_currNbParticles);
    Math::int3 grid3DRes = { subdiv3D[0], subdiv3D[1], subdiv3D[2] };
    gridVerts = Geometry::Generate3DGrid(Geometry::Shape3D::Box, grid3DRes, startFluidPos, endFluidPos, distribution);
  }
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
