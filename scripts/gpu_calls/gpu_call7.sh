set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_vector.py -m gpu -q -x -k "tensor_core" 2>&1 | tail -30 > gpurun_out/call7_tc.txt
cat gpurun_out/call7_tc.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/call7_all.txt
cat gpurun_out/call7_all.txt
timeout 300 python bench_extra.py scan --steps 10 --warmup 3 > gpurun_out/r02b_scan.jsonl 2> gpurun_out/r02b_scan.err
tail -3 gpurun_out/r02b_scan.err; cut -c1-900 gpurun_out/r02b_scan.jsonl
