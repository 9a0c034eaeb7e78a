set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/call18_tests.txt
cat gpurun_out/call18_tests.txt
timeout 300 python scripts/exp_rq.py 1000000 > gpurun_out/exp_rq4.jsonl 2> gpurun_out/exp_rq4.err
cut -c1-200 gpurun_out/exp_rq4.jsonl; tail -3 gpurun_out/exp_rq4.err
timeout 300 python bench.py --vectors 2000000 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_2M_b.json 2> gpurun_out/r02_bench_2M_b.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_2M_b.json')); print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['two_batches_in_flight']['value'], l['build'])"
timeout 900 python scripts/ef_sweep.py clustered > gpurun_out/r02_ef_sweep_clustered.jsonl 2> gpurun_out/r02_ef_sweep_clustered.err
cat gpurun_out/r02_ef_sweep_clustered.jsonl; tail -3 gpurun_out/r02_ef_sweep_clustered.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_build_1M.csv python bench_extra.py build --build-vectors 1000000 > gpurun_out/ncu_launches_build.log 2>&1
python scripts/launch_breakdown.py gpurun_out/r02_launches_build_1M.csv > gpurun_out/r02_build_1M_kernel_breakdown.txt; cat gpurun_out/r02_build_1M_kernel_breakdown.txt
