set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NIDX_B200_PRUNE=table timeout 600 python bench_extra.py build --build-vectors 1000000 > gpurun_out/r02c_build_1M_table.jsonl 2> gpurun_out/r02c_build_1M_table.err
tail -2 gpurun_out/r02c_build_1M_table.err; cut -c1-300 gpurun_out/r02c_build_1M_table.jsonl
NIDX_B200_PRUNE=table timeout 300 python -m pytest tests/test_gpu_vector.py -m gpu -q -k "build or extend" 2>&1 | tail -4
timeout 900 python bench_extra.py rabitq --build-vectors 10000000 --steps 5 --warmup 2 > gpurun_out/r02_rabitq_10M.jsonl 2> gpurun_out/r02_rabitq_10M.err
tail -2 gpurun_out/r02_rabitq_10M.err; cut -c1-700 gpurun_out/r02_rabitq_10M.jsonl
