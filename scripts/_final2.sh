mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/final_gputests.log 2>&1; tail -4 gpurun_out/final_gputests.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -c 300 gpurun_out/r02_final_bench.json
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference.json 2>> gpurun_out/r02_final_bench.err; tail -c 400 gpurun_out/r02_final_bench_reference.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:Kernel$ --profile-from-start off -f \
  -o gpurun_out/r02c_boids130k_step python scripts/step_workload.py boids_130k 320 > gpurun_out/r02c_boids.log 2>&1
tail -1 gpurun_out/r02c_boids.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
