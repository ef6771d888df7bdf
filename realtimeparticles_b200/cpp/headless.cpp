// headless.cpp -- the "new headless bench harness" of BASELINE.json's north_star in its C++ form: drives a
// Physics::Model exactly like app/ParticleSystemApp.cpp:275-318 (construct) and :377 (update() per frame), without
// SDL / OpenGL / ImGui.   usage: rtp_headless [boids|fluids|clouds] [steps] [jacobi]
#include "CudaModels.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv)
{
  const char* which = argc > 1 ? argv[1] : "fluids";
  const int steps = argc > 2 ? atoi(argv[2]) : 100;
  const int jacobi = argc > 3 ? atoi(argv[3]) : 3;
  Physics::ModelParams p;
  p.maxNbParticles = Utils::NbParticles::P130K; // ParticleSystemApp.cpp:278
  p.currNbParticles = p.maxNbParticles;
  p.dimension = Geometry::Dimension::dim3D;
  Physics::ModelType type = Physics::ModelType::FLUIDS;
  if (!strcmp(which, "boids"))
  {
    type = Physics::ModelType::BOIDS;
    p.boxSize = Geometry::BOX_SIZE_3D, p.gridRes = Geometry::GRID_RES_3D, p.pCase = Utils::PhysicsCase::BOIDS_XLARGE;
  }
  else if (!strcmp(which, "clouds"))
  {
    type = Physics::ModelType::CLOUDS;
    p.boxSize = { 10, 20, 10 }, p.gridRes = { 30, 60, 30 }, p.pCase = Utils::PhysicsCase::CLOUDS_CUMULUS; // ParticleSystemApp.cpp:300-305
  }
  else
  {
    p.boxSize = Geometry::BOX_SIZE_3D, p.gridRes = Geometry::GRID_RES_3D, p.pCase = Utils::PhysicsCase::FLUIDS_DAM;
  }
  auto model = Physics::CUDA::CreateModel(type, p);
  if (!model || !model->isInit())
  {
    fprintf(stderr, "model not initialised: %s\n", rtp_last_error(nullptr));
    return 2;
  }
  if (type != Physics::ModelType::BOIDS)
  {
    json js = model->getInputJson();
    js["Fluids"]["Nb Jacobi Iterations"][0] = jacobi;
    model->updateInputJson(js);
  }
  for (int i = 0; i < 10; ++i)
    model->update();
  rtp_handle* h = type == Physics::ModelType::BOIDS ? ((Physics::CUDA::Boids*)model.get())->handle()
      : type == Physics::ModelType::FLUIDS          ? ((Physics::CUDA::Fluids*)model.get())->handle()
                                                    : ((Physics::CUDA::Clouds*)model.get())->handle();
  rtp_sync(h);
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < steps; ++i)
    model->update();
  rtp_sync(h);
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("{\"model\": \"%s\", \"particles\": %zu, \"steps\": %d, \"steps_per_s\": %.1f, \"particle_updates_per_s\": %.4g, "
         "\"update\": \"physics + render-side kernels + camera sort (the full reference update())\"}\n",
      which, model->nbParticles(), steps, steps / s, (double)model->nbParticles() * steps / s);
  return 0;
}
