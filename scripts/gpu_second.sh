#!/bin/bash
# second GPU pass: full test suite, ncu evidence, the other workloads, reference arm
mkdir -p gpurun_out
echo "== pytest" | tee gpurun_out/second.log
timeout -k 10 1800 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread --durations=8 > gpurun_out/pytest2.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/second.log
tail -30 gpurun_out/pytest2.log
echo "== ncu launch list (cfg2)" | tee -a gpurun_out/second.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launch.err
echo "ncu launches rc=$?" | tee -a gpurun_out/second.log
echo "== ncu full (score kernel, 12.5 GB matrix)" | tee -a gpurun_out/second.log
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 2 \
    -o gpurun_out/prof_score_r01 -f python bench.py --rows 1000003 --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu_full.json 2> gpurun_out/ncu_full.err
echo "ncu full rc=$?" | tee -a gpurun_out/second.log
tail -3 gpurun_out/ncu_full.err
for wl in cfg2 cfg4 cfg3; do
  echo "== bench $wl" | tee -a gpurun_out/second.log
  extra=""
  [ "$wl" != "cfg2" ] && extra="--no-cpu-baseline"
  timeout -k 10 900 python bench.py --workload $wl --steps 10 $extra > gpurun_out/bench2_$wl.json 2> gpurun_out/bench2_$wl.err
  echo "bench $wl rc=$?" | tee -a gpurun_out/second.log
  cat gpurun_out/bench2_$wl.json; tail -3 gpurun_out/bench2_$wl.err
done
echo "== reference arm" | tee -a gpurun_out/second.log
timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench2_reference.json 2> gpurun_out/bench2_reference.err
cat gpurun_out/bench2_reference.json
nproc > gpurun_out/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt
