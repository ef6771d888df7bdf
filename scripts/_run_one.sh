N=${1:-8}; OV=${2:-1}
RTP_SLAB_OVERLAP=$OV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --slab-only > gpurun_out/slab${N}_ov$OV.json 2> gpurun_out/slab${N}_ov$OV.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/slab${N}_ov$OV.json").read().strip().splitlines()[-1])
    print("N=$N ov=$OV ms", d["ms_per_step"], "inv", d["invariants"], "phases_mid", d["phases_ms_per_rank"][len(d["phases_ms_per_rank"])//2])
except Exception as e:
    print("fail", e); print(open("gpurun_out/slab${N}_ov$OV.err").read()[-1500:])
PY
