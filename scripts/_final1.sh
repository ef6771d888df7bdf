mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/final_gputests.log 2>&1; tail -4 gpurun_out/final_gputests.log
R=r02b
B="--no-cpu-baseline --no-other-workloads"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_pbf130k.csv \
  python bench.py --steps 4 --warmup 3 $B > gpurun_out/${R}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:Kernel$ --launch-skip 3000 --launch-count 15 -f \
  -o gpurun_out/${R}_pbf130k_step python bench.py --steps 4 --warmup 200 $B > gpurun_out/${R}_full130k.log 2>&1
tail -n 2 gpurun_out/${R}_full130k.log
ls -la gpurun_out/${R}_*
