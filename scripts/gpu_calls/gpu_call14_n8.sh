set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --hybrid > gpurun_out/r02_bench_10M_n8_hybrid.json 2> gpurun_out/r02_bench_10M_n8_hybrid.err
tail -5 gpurun_out/r02_bench_10M_n8_hybrid.err
grep '^{' gpurun_out/r02_bench_10M_n8_hybrid.json | cut -c1-2500
