set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rabitq.py tests/test_gpu_zz_golden.py tests/test_gpu_mirror.py -m gpu -q 2>&1 | tail -25 > gpurun_out/call15_tests.txt
cat gpurun_out/call15_tests.txt
timeout 300 python scripts/exp_rq.py 1000000 > gpurun_out/exp_rq2.jsonl 2> gpurun_out/exp_rq2.err
cat gpurun_out/exp_rq2.jsonl; tail -3 gpurun_out/exp_rq2.err
timeout 300 python scripts/exp_hs_shape.py 2000000 > gpurun_out/exp_hs_shape3.jsonl 2> gpurun_out/exp_hs_shape3.err
cat gpurun_out/exp_hs_shape3.jsonl; tail -3 gpurun_out/exp_hs_shape3.err
timeout 600 python bench.py --vectors 2000000 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_2M_two_streams.json 2> gpurun_out/r02_bench_2M_two_streams.err
tail -3 gpurun_out/r02_bench_2M_two_streams.err; python -c "
import json; l=json.load(open('gpurun_out/r02_bench_2M_two_streams.json')); print(l['value'], l['ms_per_step'], l['two_batches_in_flight'], l['e2e'])"
