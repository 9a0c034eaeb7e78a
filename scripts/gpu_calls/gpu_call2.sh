set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_text.py tests/test_gpu_zz_golden.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/call2_text.txt
cat gpurun_out/call2_text.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/call2_pytest.txt
cat gpurun_out/call2_pytest.txt
timeout 600 python bench_extra.py bm25 --steps 5 --warmup 2 > gpurun_out/r02b_bm25.jsonl 2> gpurun_out/r02b_bm25.err
tail -3 gpurun_out/r02b_bm25.err
NIDX_B200_BM25_BITS=14 timeout 600 python bench_extra.py bm25 --steps 5 --warmup 2 > gpurun_out/r02b_bm25_bits14.jsonl 2> gpurun_out/r02b_bm25_bits14.err
cut -c1-400 gpurun_out/r02b_bm25.jsonl gpurun_out/r02b_bm25_bits14.jsonl
