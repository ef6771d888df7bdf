N=${1:-4}
run() {
  tag=$1; shift
  env RTP_SLAB_OVERLAP=1 "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --slab-only > gpurun_out/slab${N}_$tag.json 2> gpurun_out/slab${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/slab${N}_$tag.json").read().strip().splitlines()[-1])
    print("N=$N $tag ms", d["ms_per_step"], "phases_mid", d["phases_ms_per_rank"][len(d["phases_ms_per_rank"])//2])
except Exception as e:
    print("$tag fail", e); print(open("gpurun_out/slab${N}_$tag.err").read()[-800:])
PY
}
run default
run nt128 NCCL_NTHREADS=128
run nt128c1 NCCL_NTHREADS=128 NCCL_MAX_P2P_NCHANNELS=1 NCCL_MIN_P2P_NCHANNELS=1
run nt256 NCCL_NTHREADS=256
