set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/call16_tests.txt
cat gpurun_out/call16_tests.txt
timeout 300 python scripts/exp_rq.py 1000000 > gpurun_out/exp_rq3.jsonl 2> gpurun_out/exp_rq3.err
cat gpurun_out/exp_rq3.jsonl; tail -3 gpurun_out/exp_rq3.err
NIDX_B200_RQ_PREFETCH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_rabitq_kernel -c 1 -o gpurun_out/prof_rqwalk_r02b -f python bench_extra.py rabitq --steps 1 --warmup 1 > gpurun_out/ncu_rqwalk_r02b.log 2>&1
tail -2 gpurun_out/ncu_rqwalk_r02b.log | cut -c1-300
