# one GPU round trip: parity tests, then the sustained 2000-step bench (no CPU arm), summary to gpurun_out/
mkdir -p gpurun_out
T=${1:-exp}
(timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 2000 --warmup 50 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
b=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("steps/s", b["steps_per_s"], "l2res", b["l2_resident"]["steps_per_s"], "e2e", b["e2e"]["value"])
print({k:v["ms_per_launch"] for k,v in b["kernels"].items()})
for k,v in b["other_workloads"].items():
    print(k, v.get("steps_per_s"))
    if "kernels_16m" in v: print({kk:vv["ms_per_launch"] for kk,vv in v["kernels_16m"].items()})
PY
