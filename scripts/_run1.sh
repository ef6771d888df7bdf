set -x
timeout 900 python -m pytest tests/test_sharded.py -x -q -m gpu 2>&1 | tail -5
for ov in 0 1; do
RTP_SLAB_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --slab-only > gpurun_out/slab2_ov$ov.json 2> gpurun_out/slab2_ov$ov.err
tail -c 1500 gpurun_out/slab2_ov$ov.json; tail -3 gpurun_out/slab2_ov$ov.err
done
