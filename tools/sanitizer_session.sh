#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests (GPU box).  usage: bash tools/sanitizer_session.sh r02s
tag=${1:-r02s}
out=gpurun_out
mkdir -p $out
log=$out/${tag}_sanitizer.log
: > $log
run() {   # tool, then the pytest selection
  tool=$1; shift
  echo "=== $tool: $*" >> $log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest "$@" -q -m gpu -x -p no:cacheprovider >> $log 2>&1
  echo "=== exit $?" >> $log
}
# memcheck: every op / flag combination, the TMA-staged variants (metrics, fused prior), in-place gradients,
# capturable mode + CUDA graphs, the chained hyper epilogue, the real segment tables (small ones)
run memcheck tests/test_cuda_ops.py -k "not full_size and not standard_normal"
# (the captured forward/backward/step test is left out: under the sanitizer torch's autograd engine trips over its
#  own legacy-stream dependency during capture -- "operation would make the legacy stream depend on a capturing
#  blocking stream", raised inside loss.backward() -- before the sampler is reached; it passes without the tool)
run memcheck tests/test_cuda_inplace_grads_and_graphs.py -k "not captured_forward"
run memcheck tests/test_cuda_hier_priors.py -k "chained or Normal-gamma or Laplace-empirical or normal_gamma_0 or studentt_uniform_1"
run memcheck tests/test_cuda_prior_fusion.py
run memcheck tests/test_cuda_real_tables.py -k "densenet or convnet"
# racecheck: shared-memory hazards (s_red, s_chain, the staged streams behind the mbarrier)
run racecheck tests/test_cuda_ops.py -k "verlet and not full_size and not standard_normal"
run racecheck tests/test_cuda_hier_priors.py -k "chained"
run racecheck tests/test_cuda_prior_fusion.py
# initcheck: uninitialised global memory
run initcheck tests/test_cuda_inplace_grads_and_graphs.py -k "foreign or capturable_mode or pointer_table"
run initcheck tests/test_cuda_hier_priors.py -k "chained"
grep -E "^=== |passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" $log | tail -60
