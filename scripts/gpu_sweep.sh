#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout -k 10 300 python bench.py --steps 10 --no-cpu-baseline $2 > gpurun_out/benchq.json 2> gpurun_out/benchq.err
  grep '^{' gpurun_out/benchq.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1 $2: ms/step %.3f K2 %.3f ms achieved %.0f GB/s'%(d['ms_per_step'],r['kernel_ms'],r['achieved']))"
  tail -n 2 gpurun_out/benchq.err
}
for wl in "" "--emulate-shards 8"; do
run "default" "$wl"
COBSGPU_STAGES=8 run "stages=8" "$wl"
COBSGPU_STAGES=6 run "stages=6" "$wl"
COBSGPU_OCC=2 run "occ=2" "$wl"
COBSGPU_OCC=2 COBSGPU_STAGES=12 run "occ=2 stages=12" "$wl"
COBSGPU_OCC=4 run "occ=4" "$wl"
COBSGPU_OCC=1 run "occ=1" "$wl"
done
