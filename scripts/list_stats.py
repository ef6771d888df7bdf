"""Neighbour-list diagnostics along the headline run (debug aid): python scripts/list_stats.py [steps] [every]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from realtimeparticles_b200 import _abi as abi

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
every = int(sys.argv[2]) if len(sys.argv) > 2 else 10
h, _ = bench.make_pbf(abi, 0)
for i in range(steps):
    h.step(abi.STEP_PHYSICS)
    if i % every == 0 or i < 8:
        st = h.list_stats()
        n = max(st["particles"], 1)
        print(i, {k: v for k, v in st.items()}, "mean_len %.1f changed %.3f served %.3f" % (
            st["sum_len"] / n, st["cell_changed"] / n, st["cell_changed_served"] / max(st["cell_changed"], 1)), flush=True)
