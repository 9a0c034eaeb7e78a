set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_binding.py tests/test_gpu_rank_fusion.py tests/test_gpu_vector.py -m gpu -q 2>&1 | tail -4
