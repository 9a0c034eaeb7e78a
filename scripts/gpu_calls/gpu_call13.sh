set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rank_fusion.py tests/test_gpu_vector.py -m gpu -q 2>&1 | tail -25 > gpurun_out/call13_tests.txt
cat gpurun_out/call13_tests.txt
timeout 300 python bench_extra.py scan --steps 10 --warmup 3 > gpurun_out/r02c_scan.jsonl 2> gpurun_out/r02c_scan.err
tail -3 gpurun_out/r02c_scan.err; cut -c1-900 gpurun_out/r02c_scan.jsonl
timeout 300 python scripts/exp_rq.py 1000000 > gpurun_out/exp_rq.jsonl 2> gpurun_out/exp_rq.err
cat gpurun_out/exp_rq.jsonl; tail -3 gpurun_out/exp_rq.err
timeout 300 python scripts/exp_hs_shape.py 2000000 > gpurun_out/exp_hs_shape2.jsonl 2> gpurun_out/exp_hs_shape2.err
cat gpurun_out/exp_hs_shape2.jsonl; tail -3 gpurun_out/exp_hs_shape2.err
