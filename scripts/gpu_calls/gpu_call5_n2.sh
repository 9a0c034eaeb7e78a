set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_shard_nccl.py tests/test_gpu_text.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/call5_shard.txt
cat gpurun_out/call5_shard.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --vectors 2000000 > gpurun_out/r02_bench_2M_n2.json 2> gpurun_out/r02_bench_2M_n2.err
tail -5 gpurun_out/r02_bench_2M_n2.err
cut -c1-600 gpurun_out/r02_bench_2M_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --vectors 2000000 --exchange torch > gpurun_out/r02_bench_2M_n2_torch.json 2> gpurun_out/r02_bench_2M_n2_torch.err
cut -c1-300 gpurun_out/r02_bench_2M_n2_torch.json
