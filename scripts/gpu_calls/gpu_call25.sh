set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=power.limit,power.draw,temperature.gpu --format=csv
for lib in libnidx_b200_prev.so libnidx_b200.so libnidx_b200_prev.so libnidx_b200.so; do
  NIDX_B200_LIB=$lib timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/ab_$lib.json 2> gpurun_out/ab_$lib.err
  python -c "
import json; l=json.load(open('gpurun_out/ab_$lib.json')); print('$lib', round(l['value']), l['ms_per_step'], l['step_ms'], round(l['roofline']['frac'],4), l['roofline']['kernel_ms'], round(l['two_batches_in_flight']['value']), round(l['ef30']['qps']), l['clocks'], l['build']['seconds'])"
done
NIDX_B200_LIB=libnidx_b200_bm512.so timeout 600 python bench_extra.py bm25 --steps 10 --warmup 3 > gpurun_out/r02g_bm25_512.jsonl 2> gpurun_out/r02g_bm25_512.err
tail -2 gpurun_out/r02g_bm25_512.err; cut -c1-160 gpurun_out/r02g_bm25_512.jsonl
timeout 600 python bench_extra.py bm25 --steps 10 --warmup 3 > gpurun_out/r02g_bm25_256.jsonl 2> gpurun_out/r02g_bm25_256.err
cut -c1-160 gpurun_out/r02g_bm25_256.jsonl
NIDX_B200_LIB=libnidx_b200_bm512.so timeout 300 python -m pytest tests/test_gpu_text.py tests/test_gpu_zz_golden.py -m gpu -q 2>&1 | tail -3
