set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_text.py tests/test_gpu_zz_golden.py tests/test_gpu_rabitq.py -m gpu -q 2>&1 | tail -8 > gpurun_out/call17_tests.txt
cat gpurun_out/call17_tests.txt
timeout 600 python bench_extra.py bm25 --steps 10 --warmup 3 > gpurun_out/r02e_bm25.jsonl 2> gpurun_out/r02e_bm25.err
tail -2 gpurun_out/r02e_bm25.err; cut -c1-160 gpurun_out/r02e_bm25.jsonl
# ncu: the scan filter after the epilogue change, and the build kernels (1M build: every kernel once, from the middle of the build)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc_filter -s 2 -c 1 -o gpurun_out/prof_scantc_r02b -f python bench_extra.py scan --steps 2 --warmup 1 > gpurun_out/ncu_scantc_r02b.log 2>&1
tail -1 gpurun_out/ncu_scantc_r02b.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_link_kernel|reverse_link_kernel" -s 200 -c 2 -o gpurun_out/prof_build_r02 -f python bench_extra.py build --build-vectors 1000000 > gpurun_out/ncu_build_r02.log 2>&1
tail -1 gpurun_out/ncu_build_r02.log | cut -c1-200
# the headline on BASELINE.md's generators
timeout 900 python bench.py --data clustered --no-extra --cpu-seconds 5 > gpurun_out/r02_bench_10M_clustered.json 2> gpurun_out/r02_bench_10M_clustered.err
tail -2 gpurun_out/r02_bench_10M_clustered.err; cut -c1-300 gpurun_out/r02_bench_10M_clustered.json
timeout 900 python bench.py --data gauss --no-extra --no-cpu-baseline > gpurun_out/r02_bench_10M_gauss.json 2> gpurun_out/r02_bench_10M_gauss.err
tail -2 gpurun_out/r02_bench_10M_gauss.err; cut -c1-300 gpurun_out/r02_bench_10M_gauss.json
