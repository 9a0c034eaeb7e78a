set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/call20_tests.txt
cat gpurun_out/call20_tests.txt
timeout 300 python bench.py --vectors 2000000 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_2M_d.json 2> gpurun_out/r02_bench_2M_d.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_2M_d.json')); print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['two_batches_in_flight']['value'], l['build']['seconds'], l['e2e'], l['ef30'])"
