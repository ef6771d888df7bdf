"""A few steps of the 16.7M-particle PBF config (BASELINE.json configs[2]) on one GPU: the target of the 16M ncu captures.
Usage: python scripts/step16m.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from realtimeparticles_b200 import _abi as abi  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
h, _ = bench.make_workload(abi, "pbf_dam_16m_I3_vorticity_xsph", 0)
for _ in range(steps):
    h.step(abi.STEP_PHYSICS)
h.sync()
print("ran", steps, "steps of", bench.WORKLOADS["pbf_dam_16m_I3_vorticity_xsph"][0], "particles; launches per step", h.last_launch_count())
