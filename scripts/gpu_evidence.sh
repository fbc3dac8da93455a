#!/bin/bash
# 1-GPU evidence pass: full test suite, ncu launch list + full capture, all single-GPU workloads,
# reference arm, compute-sanitizer on the smoke test
mkdir -p gpurun_out
echo "== pytest"
timeout -k 10 1800 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread --durations=5 > gpurun_out/pytest_f.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_f.log
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "== sanitizer (memcheck) on smoke"
timeout -k 10 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 4 gpurun_out/sanitizer_memcheck.log
echo "== ncu launch list (cfg2)"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_f_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/ncu_f1.err
echo "== ncu full (score kernel, 12.5 GB matrix)"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 2 \
    -o gpurun_out/prof_score_f -f python bench.py --rows 1000003 --steps 2 --warmup 3 --no-cpu-baseline --no-overlap \
    > /dev/null 2> gpurun_out/ncu_f2.err
echo "ncu full rc=$?"
for wl in cfg2 cfg4 cfg3; do
  extra=""; [ "$wl" != "cfg2" ] && extra="--no-cpu-baseline"
  timeout -k 10 900 python bench.py --workload $wl --steps 20 $extra > gpurun_out/bench_f_$wl.json 2> gpurun_out/bench_f_$wl.err
  grep '^{' gpurun_out/bench_f_$wl.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl value %.4g ms/step %.3f e2e %.4g K2 %.3f ms frac %.3f launches %d phases %s cpu %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],r['kernel_ms'],r['frac'],d['gpu_launches'],d['phases_ms_per_step_rank0'], d.get('cpu_baseline',{}).get('value')))"
  tail -n 2 gpurun_out/bench_f_$wl.err
done
timeout -k 10 600 python bench.py --no-overlap --steps 20 --no-cpu-baseline > gpurun_out/bench_f_cfg2_nooverlap.json 2>/dev/null
grep '^{' gpurun_out/bench_f_cfg2_nooverlap.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg2 no-overlap value %.4g ms/step %.3f'%(d['value'],d['ms_per_step']))"
timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_f_reference.json 2> /dev/null
grep '^{' gpurun_out/bench_f_reference.json | cut -c1-200
