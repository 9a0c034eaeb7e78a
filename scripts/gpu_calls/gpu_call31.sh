set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_vector.py tests/test_gpu_rabitq.py tests/test_gpu_mirror.py -m gpu -q 2>&1 | tail -5
