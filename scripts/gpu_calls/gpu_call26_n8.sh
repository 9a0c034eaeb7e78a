set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/call26_tests.txt
cat gpurun_out/call26_tests.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_10M_n8_final.json 2> gpurun_out/r02_bench_10M_n8_final.err
tail -3 gpurun_out/r02_bench_10M_n8_final.err | cut -c1-300
grep '^{' gpurun_out/r02_bench_10M_n8_final.json | cut -c1-1200
