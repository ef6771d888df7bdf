# ncu --set full of every kernel of one settled boids step and one clouds step (through gpurun); summarised into profiles/ by scripts/ncu_summary.py.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sharded.py -m gpu -q -x -k "library" 2>&1 | tail -2
# boids: the 7 launches of step 320 (graph replays are not profiled: only the last, plain step is)
timeout 250 ncu --set full --clock-control none --import-source on -k regex:Kernel$ --profile-from-start off -f \
  -o gpurun_out/r02b_boids130k_step python scripts/step_workload.py boids_130k 320 > gpurun_out/r02b_boids.log 2>&1
tail -2 gpurun_out/r02b_boids.log
timeout 250 ncu --set full --clock-control none --import-source on -k regex:Kernel$ --profile-from-start off -f \
  -o gpurun_out/r02b_clouds130k_step python scripts/step_workload.py clouds_130k_I2 200 > gpurun_out/r02b_clouds.log 2>&1
tail -2 gpurun_out/r02b_clouds.log
ls -la gpurun_out/r02b_boids* gpurun_out/r02b_clouds*
