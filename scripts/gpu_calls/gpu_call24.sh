set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vector.py -m gpu -q -k "alternating" 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_10M_final.json 2> gpurun_out/r02_bench_10M_final.err
tail -3 gpurun_out/r02_bench_10M_final.err; cut -c1-400 gpurun_out/r02_bench_10M_final.json
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_10M_reference_arm.json 2> gpurun_out/r02_bench_10M_reference_arm.err
tail -2 gpurun_out/r02_bench_10M_reference_arm.err; cut -c1-600 gpurun_out/r02_bench_10M_reference_arm.json
