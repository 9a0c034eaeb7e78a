set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_text.py tests/test_gpu_zz_golden.py tests/test_gpu_mirror.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/call4_text.txt
cat gpurun_out/call4_text.txt
timeout 600 python bench_extra.py bm25 --steps 10 --warmup 3 > gpurun_out/r02c_bm25.jsonl 2> gpurun_out/r02c_bm25.err
tail -3 gpurun_out/r02c_bm25.err
cut -c1-300 gpurun_out/r02c_bm25.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_kernel -s 1 -c 1 -o gpurun_out/prof_bm25_r02c -f python bench_extra.py bm25 --steps 2 --warmup 1 > gpurun_out/ncu_bm25_r02c.log 2>&1
tail -3 gpurun_out/ncu_bm25_r02c.log | cut -c1-300
