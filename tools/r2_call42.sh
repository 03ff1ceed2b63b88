#!/bin/bash
# round 2, call 42: the whole GPU suite on the final build (4 workers; anything that fails is re-run serially)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m pytest tests -m gpu -q -n 4 --timeout 600 > gpurun_out/c42_suite.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c42_suite.log | cut -c1-300
if ! tail -1 gpurun_out/c42_suite.log | grep -q " passed" || tail -1 gpurun_out/c42_suite.log | grep -q "failed\|error"; then
  timeout -k 10 600 python -m pytest tests -m gpu -q --lf --timeout 600 > gpurun_out/c42_suite_lf.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/c42_suite_lf.log | cut -c1-300
fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c42_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/c42_smoke.log
