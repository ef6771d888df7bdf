"""Run a few un-graphed steps of one BASELINE.json workload (for ncu): python scripts/step_workload.py boids_130k 320
(the kernels of the LAST step are the ones a --launch-skip picks)."""
import sys
sys.path.insert(0, ".")
import bench
from realtimeparticles_b200 import _abi as abi
name, steps = sys.argv[1], int(sys.argv[2])
h, _ = bench.make_workload(abi, name, 0)
h.step_n(steps - 1, abi.STEP_PHYSICS)  # one graph
h.sync()
import torch
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the last step is profiled
h.step(abi.STEP_PHYSICS)
h.sync()
torch.cuda.profiler.stop()
print("launches per step", h.last_launch_count())
