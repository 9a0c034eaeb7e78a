set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rabitq.py tests/test_gpu_rank_fusion.py tests/test_gpu_zz_golden.py -m gpu -q 2>&1 | tail -5
timeout 300 python scripts/exp_rq.py 1000000 > gpurun_out/exp_rq6.jsonl 2> gpurun_out/exp_rq6.err
cut -c1-170 gpurun_out/exp_rq6.jsonl; tail -3 gpurun_out/exp_rq6.err
