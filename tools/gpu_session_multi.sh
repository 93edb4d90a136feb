#!/bin/bash
# Multi-GPU session on one box: NCCL parity test, smoke(), bench.py under torchrun at N = 1, 2, 4, 8 (as far as
# the box goes).   usage: bash tools/gpu_session_multi.sh r02m
tag=${1:-r02m}
out=gpurun_out
mkdir -p $out
ngpu=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
python -m pytest tests/test_cuda_chains_nccl.py tests/test_cuda_example_chains.py -q -m gpu -p no:cacheprovider > $out/${tag}_nccl_tests.log 2>&1; tail -3 $out/${tag}_nccl_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
port=29600
for n in 1 2 4 8; do
  [ $n -le $ngpu ] || continue
  for steps in 20 200; do
    # the long run only at the ends of the range (an 8-GPU box is charged eight-fold)
    if [ $steps -eq 200 ] && [ $n -ne 1 ] && [ $n -ne $ngpu ]; then continue; fi
    port=$((port+1))
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps $steps --warmup 5 --no-extra --no-cpu > $out/${tag}_bench_n${n}_s${steps}.json 2>> $out/${tag}_bench.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --steps $steps --warmup 5 --no-extra --no-cpu > $out/${tag}_bench_n${n}_s${steps}.json 2>> $out/${tag}_bench.err
    fi
    python - <<PY
import json
d = json.load(open("$out/${tag}_bench_n${n}_s${steps}.json"))
print("N=$n steps=$steps value %.4g ms/step %.4f kernel_us %.2f prod_us %.2f e2e %.4g h2d_GBs %.1f gather_ms %s host_us %.1f" % (
    d["value"], d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["production"]["kernel_us"], d["e2e"]["value"],
    d["e2e"]["bare_h2d_GBs_per_gpu"], d.get("cycle_gather_ms"), d["impl_notes"]["host_us_per_step"]),
    {k: round(v["value"] / 1e11, 3) for k, v in d["samplers"].items()})
PY
  done
done
tail -5 $out/${tag}_bench.err
