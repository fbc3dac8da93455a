#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -x > gpurun_out/pytest3.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest3.log
echo "== ncu launch list (cfg2)"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches3_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/ncu3.err
echo "== bench cfg2"
timeout -k 10 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench3_cfg2.json 2> gpurun_out/bench3_cfg2.err
grep '^{' gpurun_out/bench3_cfg2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4g ms/step %.3f e2e %.4g K2 %.3f ms frac %.3f launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],r['kernel_ms'],r['frac'],d['gpu_launches']))"
tail -n 3 gpurun_out/bench3_cfg2.err
