#!/bin/bash
# One GPU-box session: tests, bench (both arms), small-model latencies, ncu evidence.  Everything lands in
# gpurun_out/ with the tag given as $1.   usage: bash tools/gpu_session.sh r02c [skip-tests]
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $out/${tag}_clocks.csv 2>/dev/null &
smi=$!
if [ "$2" != "skip-tests" ]; then
  BNNP_REPORT_FILE=$PWD/$out/${tag}_parity_reports.jsonl python -m pytest tests -q -m gpu -p no:cacheprovider > $out/${tag}_gpu_tests.log 2>&1
  tail -25 $out/${tag}_gpu_tests.log
fi
python bench.py --steps 200 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || tail -20 $out/${tag}_bench.err
python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > $out/${tag}_bench_20steps.json 2>> $out/${tag}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err || tail -5 $out/${tag}_bench_ref.err
python tools/bench_small_models.py > $out/${tag}_small_models.json 2> $out/${tag}_small_models.err || tail -5 $out/${tag}_small_models.err
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
python tools/variant_times.py > $out/${tag}_variants.jsonl 2>> $out/${tag}_small_models.err
python tools/tune_tiles.py > $out/${tag}_tune.json 2>> $out/${tag}_small_models.err
kill $smi 2>/dev/null
# ---- ncu: launch list of the bench command (shares of the step), then the kernels themselves
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > $out/${tag}_ncu_bench.log 2>&1
# the step kernel with the caches flushed before every replay = the production regime
ncu --set full --clock-control none --import-source on -k regex:bnnp_step_kernel -s 4 -c 2 -f -o $out/${tag}_step_cold \
    python tools/ncu_target.py sgld > $out/${tag}_ncu_cold.log 2>&1
# consecutive launches, one pass, caches as the previous launch left them = the back-to-back regime
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -k regex:bnnp_step_kernel -s 8 -c 16 --csv --log-file $out/${tag}_dram_steady.csv python tools/ncu_target.py sgld 30 > /dev/null 2>&1
# gradients read in place (tensors of their own): DRAM bytes and the launch list (no multi-tensor copy)
ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -s 40 -c 60 --csv --log-file $out/${tag}_foreign_launches.csv python tools/ncu_target.py sgld_foreign 16 > /dev/null 2>&1
ncu -i $out/${tag}_step_cold.ncu-rep --page raw --csv > $out/${tag}_step_cold_raw.csv 2>/dev/null
for c in verlet_fused sgld_metrics hmc verlet_save; do
  ncu --set full --clock-control none -k regex:bnnp_step_kernel -s 4 -c 1 -f -o $out/${tag}_${c}_cold \
      python tools/ncu_target.py $c > /dev/null 2>&1
  # gpurun brings back at most 64 MiB: keep the raw metric page of the variants, the report only of the main kernel
  ncu -i $out/${tag}_${c}_cold.ncu-rep --page raw --csv > $out/${tag}_${c}_cold_raw.csv 2>/dev/null
  rm -f $out/${tag}_${c}_cold.ncu-rep
done
du -sh $out; ls -la $out | tail -30
