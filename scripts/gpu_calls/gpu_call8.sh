set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/call8_pytest.txt
cat gpurun_out/call8_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/call8_smoke.txt 2>&1; tail -3 gpurun_out/call8_smoke.txt
timeout 900 python bench.py > gpurun_out/r02_bench_10M.json 2> gpurun_out/r02_bench_10M.err
tail -5 gpurun_out/r02_bench_10M.err
cut -c1-1500 gpurun_out/r02_bench_10M.json
