set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rabitq.py tests/test_gpu_zz_golden.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/call10_rabitq.txt
cat gpurun_out/call10_rabitq.txt
timeout 300 python bench_extra.py rabitq --steps 5 --warmup 2 > gpurun_out/r02b_rabitq.jsonl 2> gpurun_out/r02b_rabitq.err
tail -3 gpurun_out/r02b_rabitq.err; cut -c1-400 gpurun_out/r02b_rabitq.jsonl
# ncu: BM25 (OR-50 is the first bm25_kernel instantiation launched by bench_extra bm25)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_kernel -s 3 -c 1 -o gpurun_out/prof_bm25_r02 -f python bench_extra.py bm25 --steps 2 --warmup 1 > gpurun_out/ncu_bm25_r02.log 2>&1
tail -3 gpurun_out/ncu_bm25_r02.log | cut -c1-300
# ncu: the headline kernel at 1M (same per-query work as 10M up to log N) + launch list of the timed region
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel --profile-from-start off -c 1 -o gpurun_out/prof_hnsw_r02 -f python bench.py --vectors 2000000 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_hnsw_r02.log 2>&1
tail -3 gpurun_out/ncu_hnsw_r02.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_bench_2M.csv python bench.py --vectors 2000000 --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches_r02.log 2>&1
tail -5 gpurun_out/r02_launches_bench_2M.csv
ls -la gpurun_out
