#!/bin/bash
# One GPU-box visit: GPU parity tests, smoke, bench (run under gpurun from the repo root).  $1 = tag for the logs.
T=${1:-rX}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -vE "^\s*$" > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.log 2>&1; tail -c 2500 gpurun_out/${T}_bench.log
