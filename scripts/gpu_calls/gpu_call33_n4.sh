set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_10M_n4.json 2> gpurun_out/r02_bench_10M_n4.err
tail -2 gpurun_out/r02_bench_10M_n4.err | cut -c1-200
grep '^{' gpurun_out/r02_bench_10M_n4.json | cut -c1-400
