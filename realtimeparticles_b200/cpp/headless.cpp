// headless.cpp -- the "new headless bench harness" of BASELINE.json's north_star in its C++ form: drives a
// Physics::Model exactly like app/ParticleSystemApp.cpp:275-318 (construct) and :377 (update() per frame), without
// SDL / OpenGL / ImGui.   usage: rtp_headless [boids|fluids|clouds] [steps] [jacobi]
//                           rtp_headless slabs <nslabs> [steps] [warmup]   -- BASELINE.json configs[4]: the 16.7M-particle PBF
//                           dam break, x-slab decomposed over the GPUs of the box by rtp_slab_group (one host thread)
#include "CudaModels.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// the 16.7M-particle dam (box 80x40x40, grid 240x120x120) on `nslabs` slabs spread over the visible GPUs
static int runSlabs(int nslabs, int steps, int warmup)
{
  const uint32_t box[3] = { 80, 40, 40 }, grid[3] = { 240, 120, 120 };
  const int res[3] = { 512, 256, 128 };
  const float start[3] = { -40.0f, -20.0f, -20.0f }, end[3] = { 40.0f, 0.0f, 0.0f };
  const uint64_t total = (uint64_t)res[0] * res[1] * res[2];
  std::vector<float> pos(total * 4), vel(total * 4, 0.0f);
  if (rtp_gen_box_grid(pos.data(), res, start, end) != (int64_t)total)
    return 2;
  const int ndev = rtp_device_count();
  if (ndev < 1 || nslabs < 1)
  {
    fprintf(stderr, "no CUDA device\n");
    return 2;
  }
  std::vector<int> devs(nslabs);
  for (int r = 0; r < nslabs; ++r) // GPUs 0 .. nslabs-1; with more slabs than GPUs, contiguous slabs share one
    devs[r] = nslabs <= ndev ? r : (int)((long long)r * ndev / nslabs) % ndev;
  // capacities of bench.py's run_slab_16m: ghost regions of 2 x-layers with room for 8 particles per cell, 32k-row migration messages
  const uint64_t ghostCap = nslabs > 1 ? 2ull * 120 * 120 * 8 : 0, migrateCap = nslabs > 1 ? (1ull << 15) : 1;
  const uint64_t capacity = (uint64_t)((double)(total / nslabs) * 1.10) + 2 * migrateCap + 2 * ghostCap;
  rtp_slab_group* g = nullptr;
  if (rtp_slab_group_create(&g, nslabs, devs.data(), capacity, box, grid, ghostCap, migrateCap, 1) != RTP_OK)
  {
    fprintf(stderr, "rtp_slab_group_create: %s\n", rtp_slab_group_last_error(nullptr));
    return 2;
  }
  const rtp_fluid_params fp = { 450.0f, 600.0f, 0.010f, 3, 1, 0.006f, 0.001f, 4, 1, 0.0004f, 0.0001f };
  int rc = rtp_slab_group_set_fluid_params(g, &fp, 3);
  if (rc == RTP_OK)
    rc = rtp_slab_group_upload(g, pos.data(), vel.data(), total);
  if (rc == RTP_OK)
    rc = rtp_slab_group_step(g, warmup);
  if (rc == RTP_OK)
    rc = rtp_slab_group_sync(g);
  const auto t0 = std::chrono::steady_clock::now();
  if (rc == RTP_OK)
    rc = rtp_slab_group_step(g, steps);
  const double enqueue = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc == RTP_OK)
    rc = rtp_slab_group_sync(g);
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t migrated = 0;
  if (rc == RTP_OK)
    rc = rtp_slab_group_check(g, &migrated);
  int64_t n = -1;
  std::vector<uint64_t> per(nslabs);
  if (rc == RTP_OK)
    n = rtp_slab_group_download(g, pos.data(), vel.data(), total, per.data());
  if (rc != RTP_OK || n < 0)
  {
    fprintf(stderr, "slab group: %s\n", rtp_slab_group_last_error(g));
    rtp_slab_group_destroy(g);
    return 2;
  }
  double ke = 0.0;
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k)
      ke += 0.5 * (double)vel[4 * i + k] * (double)vel[4 * i + k];
  printf("{\"workload\": \"pbf_dam_16m_I3_vorticity_xsph, rtp_slab_group\", \"slabs\": %d, \"gpus\": %d, \"particles\": %lld, \"steps\": %d, \"warmup\": %d, "
         "\"ms_per_step\": %.3f, \"host_enqueue_ms_per_step\": %.3f, \"particle_updates_per_s\": %.4g, \"migrated_total\": %llu, "
         "\"kinetic_energy\": %.9g, \"after_step\": %d, \"timing\": \"host wall clock around enqueue + device drain\"}\n",
      nslabs, ndev < nslabs ? ndev : nslabs, (long long)n, steps, warmup, s * 1e3 / steps, enqueue * 1e3 / steps, (double)n * steps / s,
      (unsigned long long)migrated, ke, warmup + steps);
  rtp_slab_group_destroy(g);
  return 0;
}

int main(int argc, char** argv)
{
  const char* which = argc > 1 ? argv[1] : "fluids";
  if (!strcmp(which, "slabs"))
    return runSlabs(argc > 2 ? atoi(argv[2]) : 1, argc > 3 ? atoi(argv[3]) : 10, argc > 4 ? atoi(argv[4]) : 9);
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
