set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_vector.py -m gpu -q -x -k "tensor_core" 2>&1 | tail -30 > gpurun_out/call6_tc.txt
echo "rc=$?" >> gpurun_out/call6_tc.txt
cat gpurun_out/call6_tc.txt
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
timeout 300 python -m pytest tests/test_gpu_text.py tests/test_gpu_vector.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/call6_rest.txt
cat gpurun_out/call6_rest.txt
timeout 300 python bench_extra.py scan --steps 10 --warmup 3 > gpurun_out/r02_scan.jsonl 2> gpurun_out/r02_scan.err
tail -3 gpurun_out/r02_scan.err; cut -c1-700 gpurun_out/r02_scan.jsonl
timeout 600 python bench_extra.py bm25 --steps 10 --warmup 3 > gpurun_out/r02d_bm25.jsonl 2> gpurun_out/r02d_bm25.err
tail -3 gpurun_out/r02d_bm25.err; cut -c1-200 gpurun_out/r02d_bm25.jsonl
