set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/call19_tests.txt
cat gpurun_out/call19_tests.txt
timeout 600 python bench_extra.py build --build-vectors 1000000 > gpurun_out/r02_build_1M.jsonl 2> gpurun_out/r02_build_1M.err
tail -2 gpurun_out/r02_build_1M.err; cut -c1-420 gpurun_out/r02_build_1M.jsonl
timeout 300 python scripts/exp_rq.py 1000000 > gpurun_out/exp_rq5.jsonl 2> gpurun_out/exp_rq5.err
cut -c1-170 gpurun_out/exp_rq5.jsonl; tail -3 gpurun_out/exp_rq5.err
timeout 300 python bench.py --vectors 2000000 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_2M_c.json 2> gpurun_out/r02_bench_2M_c.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_2M_c.json')); print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['two_batches_in_flight']['value'], l['build'], l['e2e'])"
