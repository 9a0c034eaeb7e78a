set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_10M_final2.json 2> gpurun_out/r02_bench_10M_final2.err
tail -2 gpurun_out/r02_bench_10M_final2.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_10M_final2.json')); print(l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['kernel_ms'], l['roofline']['kernel_ms_accounting_pass'], l['e2e'], l['clocks'], [ (e['config']['workload'][:40], round(e['value'])) for k in l['extra'] for e in l['extra'][k]])"
