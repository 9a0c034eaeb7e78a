set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_vector.py -m gpu -q -x -k "tensor_core" 2>&1 | tail -12
timeout 200 python bench_extra.py scan --steps 10 --warmup 3 > gpurun_out/r02e_scan.jsonl 2> gpurun_out/r02e_scan.err
tail -3 gpurun_out/r02e_scan.err; cut -c1-620 gpurun_out/r02e_scan.jsonl | head -1
