set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rank_fusion.py tests/test_gpu_vector.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/call11_tests.txt
cat gpurun_out/call11_tests.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_rabitq_kernel -c 1 -o gpurun_out/prof_rqwalk_r02 -f python bench_extra.py rabitq --steps 1 --warmup 1 > gpurun_out/ncu_rqwalk_r02.log 2>&1
tail -2 gpurun_out/ncu_rqwalk_r02.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc_ -s 2 -c 2 -o gpurun_out/prof_scantc_r02 -f python bench_extra.py scan --steps 2 --warmup 1 > gpurun_out/ncu_scantc_r02.log 2>&1
tail -2 gpurun_out/ncu_scantc_r02.log | cut -c1-300
ls -la gpurun_out
