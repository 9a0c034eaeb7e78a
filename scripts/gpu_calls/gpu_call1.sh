set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
scripts/bin/ubench_smem > gpurun_out/ubench_smem.txt 2>&1
NIDX_B200_UNVERIFIED_GPU_TESTS=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/call1_pytest.txt
cat gpurun_out/call1_pytest.txt
for v in default w128 w256; do
  if [ $v = default ]; then unset NIDX_B200_BM25; else export NIDX_B200_BM25=$v; fi
  timeout 600 python bench_extra.py bm25 --steps 5 --warmup 2 > gpurun_out/r02_bm25_$v.jsonl 2> gpurun_out/r02_bm25_$v.err
done
tail -3 gpurun_out/ubench_smem.txt
