set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --vectors 2000000 --no-extra --cpu-seconds 3 > gpurun_out/r02_bench_2M_e.json 2> gpurun_out/r02_bench_2M_e.err
tail -2 gpurun_out/r02_bench_2M_e.err
python -c "
import json; l=json.load(open('gpurun_out/r02_bench_2M_e.json')); print(l['value'], l['ms_per_step'], l['step_ms'], l['roofline'], l['clocks'])"
