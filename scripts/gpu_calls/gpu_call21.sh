set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/call21_tests.txt
cat gpurun_out/call21_tests.txt
timeout 600 python bench_extra.py bm25 --steps 10 --warmup 3 > gpurun_out/r02f_bm25.jsonl 2> gpurun_out/r02f_bm25.err
tail -2 gpurun_out/r02f_bm25.err; cut -c1-160 gpurun_out/r02f_bm25.jsonl
timeout 600 python bench_extra.py build --build-vectors 1000000 > gpurun_out/r02b_build_1M.jsonl 2> gpurun_out/r02b_build_1M.err
tail -2 gpurun_out/r02b_build_1M.err; cut -c1-300 gpurun_out/r02b_build_1M.jsonl
