#!/usr/bin/env bash
# What the round-end driver does, in one place.
#   scripts/round_check.sh cpu    here (no GPU): build everything, CPU test suite
#   scripts/round_check.sh gpu    on a B200 box (e.g. `gpurun --timeout 900 -- scripts/round_check.sh gpu`): GPU suite, smoke,
#                                 a short bench (1 M vectors) with its ncu launch list under gpurun_out/
set -euo pipefail
cd "$(dirname "$0")/.."
case "${1:-cpu}" in
cpu)
    python -c "import __graft_entry__ as g; g.build()"
    python -m pytest tests -x -q -m "not gpu"
    python scripts/dry_run_gpu_golden.py
    ;;
gpu)
    mkdir -p gpurun_out
    NIDX_B200_UNVERIFIED_GPU_TESTS="${NIDX_B200_UNVERIFIED_GPU_TESTS:-0}" python -m pytest tests -q -m gpu
    python -c "import __graft_entry__ as g; g.smoke()"
    python bench.py --vectors 1000000 --steps 20 --warmup 3 | tee gpurun_out/bench_1M.json
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_1M.csv \
        python bench.py --vectors 1000000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1 || true
    ;;
*)
    echo "usage: $0 cpu|gpu" >&2
    exit 2
    ;;
esac
