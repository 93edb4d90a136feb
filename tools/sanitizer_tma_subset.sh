#!/bin/bash
# compute-sanitizer over the tests that run the TMA-staged kernel variants (fused prior, metrics): the short
# re-check after a change of that path; tools/sanitizer_session.sh is the full session.  GPU box only.
out=gpurun_out; log=$out/r02z_sanitizer.log; : > $log
run() { tool=$1; shift; echo "=== $tool: $*" >> $log; timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest "$@" -q -m gpu -x -p no:cacheprovider >> $log 2>&1; echo "=== exit $?" >> $log; }
run memcheck tests/test_cuda_prior_fusion.py
run racecheck tests/test_cuda_prior_fusion.py
run racecheck tests/test_cuda_ops.py -k "verlet and not full_size and not standard_normal"
grep -E "^=== |passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" $log | tail -30
