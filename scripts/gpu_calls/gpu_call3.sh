set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rabitq.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/call3_rabitq.txt
cat gpurun_out/call3_rabitq.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_kernel -s 1 -c 1 -o gpurun_out/prof_bm25_r02 -f python bench_extra.py bm25 --steps 2 --warmup 1 > gpurun_out/ncu_bm25_r02.log 2>&1
tail -5 gpurun_out/ncu_bm25_r02.log
