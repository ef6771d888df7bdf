N=${1:-8}
for ov in 0 1; do
RTP_SLAB_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --slab-only > gpurun_out/slab${N}_ov$ov.json 2> gpurun_out/slab${N}_ov$ov.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/slab${N}_ov$ov.json").read().strip().splitlines()[-1])
    print("N=$N ov=$ov ms", d["ms_per_step"], "inv", d["invariants"], "phases0", d["phases_ms_per_rank"][0], "phases_mid", d["phases_ms_per_rank"][len(d["phases_ms_per_rank"])//2])
except Exception as e:
    print("fail", e); print(open("gpurun_out/slab${N}_ov$ov.err").read()[-1500:])
PY
done
