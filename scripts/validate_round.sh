# Round-end validation on a GPU box (through gpurun): the -m gpu tests, smoke(), the default bench line, the C++ headless harness.
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/final_gputests.log 2>&1; tail -5 gpurun_out/final_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -c 200 gpurun_out/r02_final_bench.json; tail -2 gpurun_out/r02_final_bench.err
(cd realtimeparticles_b200/lib && timeout 60 ./rtp_headless fluids 300 3)
