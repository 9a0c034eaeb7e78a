set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_shard_nccl.py tests/test_gpu_vector.py -m gpu -q -x -k "shard or nccl or normalize" 2>&1 | tail -15 > gpurun_out/call9_shard.txt
cat gpurun_out/call9_shard.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --vectors 1000000 --hybrid > gpurun_out/r02_bench_1M_n2_hybrid.json 2> gpurun_out/r02_bench_1M_n2_hybrid.err
tail -5 gpurun_out/r02_bench_1M_n2_hybrid.err
cut -c1-3000 gpurun_out/r02_bench_1M_n2_hybrid.json
