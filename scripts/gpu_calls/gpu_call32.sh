set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel --profile-from-start off -c 1 -o gpurun_out/prof_hnsw_10M_r02 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_hnsw_10M_r02.log 2>&1
tail -2 gpurun_out/ncu_hnsw_10M_r02.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_bench_10M.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches_10M_r02.log 2>&1
tail -4 gpurun_out/r02_launches_bench_10M.csv | cut -c1-200
