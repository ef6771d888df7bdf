// CudaModels.hpp -- the CUDA backend behind the reference's UNMODIFIED physics interface.
//
//   Physics::CUDA::Boids / Fluids / Clouds derive from Physics::Model (physics/Model.hpp:81-216) exactly like
//   Physics::CL::Boids / Fluids / Clouds do (physics/ocl/{Boids,Fluids,Clouds}.hpp), own the same JSON parameter
//   schema and the same POD kernel-input structs, and forward every device operation to the C ABI of
//   include/rtp_cuda.h (librtp_cuda.so). A maintainer switches backend by calling Physics::CUDA::CreateModel
//   instead of the factory body in physics/Model.cpp:10-24.
//
// Compiled against the reference's own headers (-I<ref>/physics -I<ref>/utils); nothing of the reference is copied.
#pragma once

#include "Model.hpp" // the reference's physics/Model.hpp

#include "../../include/rtp_cuda.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <cstdlib>
#include <utility>
#include <variant>
#include <vector>

namespace Physics
{
class Target; // physics/utils/Target.hpp
}

namespace Physics::CUDA
{
// same field order / size as the reference's kernel-input PODs (Boids.hpp:14-26, Fluids.hpp:17-32, Clouds.hpp:15-45)
using BoidsRuleKernelInputs = rtp_boids_params;
using TargetKernelInputs = rtp_target_params;
using FluidKernelInputs = rtp_fluid_params;
using CloudKernelInputs = rtp_cloud_params;

// which GPU the backend runs on: RTP_DEVICE in the environment (default 0); the OpenCL backend picks "the" GPU itself
inline int deviceFromEnv()
{
  const char* e = std::getenv("RTP_DEVICE");
  return e ? std::atoi(e) : 0;
}

// CUDA counterpart of OclModel<KernelInputs...> (physics/ocl/OclModel.hpp:12-70)
template <typename... KernelInputs>
class CudaModel : public Model
{
  public:
  CudaModel(ModelParams params, int rtpModel, unsigned maxPartsInCell, KernelInputs... kernelInputs, json inputJson = {})
      : Model(params, inputJson)
  {
    (m_kernelInputs.push_back(kernelInputs), ...);
    rtp_config cfg {};
    cfg.model = rtpModel;
    cfg.device = deviceFromEnv(); // RTP_DEVICE selects the GPU (default 0); the reference has no such notion
    cfg.max_particles = params.maxNbParticles;
    cfg.nb_particles = params.currNbParticles <= params.maxNbParticles ? params.currNbParticles : params.maxNbParticles;
    cfg.box[0] = (uint32_t)params.boxSize.x, cfg.box[1] = (uint32_t)params.boxSize.y, cfg.box[2] = (uint32_t)params.boxSize.z;
    cfg.grid[0] = (uint32_t)params.gridRes.x, cfg.grid[1] = (uint32_t)params.gridRes.y, cfg.grid[2] = (uint32_t)params.gridRes.z;
    cfg.dim = params.dimension == Geometry::Dimension::dim2D ? 2 : 3;
    cfg.max_parts_in_cell = maxPartsInCell;
    if (rtp_create(&cfg, &m_handle) != RTP_OK)
    {
      // like a failed OpenCL context: the model stays un-initialised and the app shows its pop-up
      m_handle = nullptr;
      m_createError = rtp_last_error(nullptr);
    }
    // the VBOs of the render engine (Model.hpp:67-70, created by render/Engine.cpp:70-87, :300-304) become the buffers of
    // p_pos / p_col / c_partDetector, like the reference's createGLBuffer calls (Fluids.cpp:132-135). 0 = headless host.
    if (m_handle)
    {
      const std::pair<int, unsigned int> shared[] = { { RTP_F_POS, params.particlePosVBO }, { RTP_F_COL, params.particleColVBO },
        { RTP_F_PART_DETECTOR, params.gridVBO } };
      for (const auto& sh : shared)
        if (sh.second != 0 && rtp_register_gl(m_handle, sh.first, sh.second) != RTP_OK)
          m_createError = rtp_last_error(m_handle); // keeps running on the library's own buffer for this field
    }
  }
  ~CudaModel() override { rtp_destroy(m_handle); }

  bool isProfilingEnabled() const override { return m_profiling; }
  void enableProfiling(bool enable) override
  {
    m_profiling = enable;
    if (m_handle)
      rtp_enable_profiling(m_handle, enable ? 1 : 0);
  }
  bool isUsingIGPU() const override { return false; }

  void updateModelWithInputJson(json& inputJson) override
  {
    transferJsonInputsToModel(inputJson);
    transferKernelInputsToGPU();
  }
  virtual void transferJsonInputsToModel(json& inputJson) = 0;
  virtual void transferKernelInputsToGPU() = 0;

  template <typename T>
  T& getKernelInput(int index) { return std::get<T>(m_kernelInputs.at(index)); }
  size_t getNbKernelInputs() { return m_kernelInputs.size(); }

  // harness-facing additions (the app reaches the same data through the shared GL buffers)
  rtp_handle* handle() { return m_handle; }
  const std::string& createError() const { return m_createError; }
  void setCameraPos(const Math::float3& cam) { m_camera[0] = cam.x, m_camera[1] = cam.y, m_camera[2] = cam.z; }
  void setStepFlags(unsigned flags) { m_stepFlags = flags; }

  protected:
  // replayGraph: the frame is the cached CUDA graph of the step (one launch, same result; the library falls back to plain
  // launches while OpenGL buffers are registered). Not when profiling (events between the stages) or when the kernel
  // parameters change every frame (the moving boids target).
  void stepDevice(bool replayGraph = true)
  {
    unsigned flags = m_stepFlags;
    if (m_pause)
      flags &= ~RTP_STEP_PHYSICS; // Fluids.cpp:409: on pause only the camera sort (and clouds colouring) run
    if (replayGraph && !m_profiling)
      rtp_step_n(m_handle, flags, m_camera, 1);
    else
      rtp_step(m_handle, flags, m_camera);
  }
  void uploadParticles(const std::vector<Math::float3>& verts, bool velocityIsPosition, const float* colour);

  std::vector<std::variant<KernelInputs...>> m_kernelInputs;
  rtp_handle* m_handle = nullptr;
  std::string m_createError;
  bool m_profiling = false;
  float m_camera[3] = { 32.0f, -1.2f, 0.0f }; // render/Camera.cpp:11
  unsigned m_stepFlags = RTP_STEP_PHYSICS | RTP_STEP_RENDER_AUX | RTP_STEP_CAMERA_SORT;
};

class Boids : public CudaModel<BoidsRuleKernelInputs, TargetKernelInputs>
{
  public:
  Boids(ModelParams params);
  ~Boids() override;
  void update() override;
  void reset() override;
  Math::float3 targetPos() const override;
  bool isTargetActivated() const override;
  bool isTargetVisible() const override;
  void transferJsonInputsToModel(json& inputJson) override;
  void transferKernelInputsToGPU() override;

  private:
  void initBoidsParticles();
  std::unique_ptr<Target> m_target;
};

class Fluids : public CudaModel<FluidKernelInputs>
{
  public:
  Fluids(ModelParams params);
  void update() override;
  void reset() override;
  void transferJsonInputsToModel(json& inputJson) override;
  void transferKernelInputsToGPU() override;

  private:
  void initFluidsParticles();
  size_t m_nbJacobiIters = 2;
};

class Clouds : public CudaModel<FluidKernelInputs, CloudKernelInputs>
{
  public:
  Clouds(ModelParams params);
  void update() override;
  void reset() override;
  void transferJsonInputsToModel(json& inputJson) override;
  void transferKernelInputsToGPU() override;

  private:
  void initCloudsParticles();
  void pushDisplayedQuantity();
  size_t m_nbJacobiIters = 1;
};

// drop-in for Physics::CreateModel (physics/Model.cpp:10-24)
std::unique_ptr<Model> CreateModel(ModelType type, ModelParams params);
}
