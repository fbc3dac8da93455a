#!/bin/bash
mkdir -p gpurun_out
show() { grep '^{' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$2 N=%d value %.4g ms/step %.3f e2e %.4g K2 %.3f ms frac %.3f wsf %.3f clocks %s'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['kernel_ms'],r['frac'],r['whole_step_frac'],d['clocks']['reasons']))"; }
run() { # name N args...
  name=$1; N=$2; shift 2
  if [ "$N" = "1" ]; then
    timeout -k 10 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/sf_$name.json 2> gpurun_out/sf_$name.err
  else
    timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
      bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/sf_$name.json 2> gpurun_out/sf_$name.err
  fi
  echo "rc=$?"; show gpurun_out/sf_$name.json $name
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/sf_$name.err | tail -n 3
}
run cfg2_x1 1
run cfg2_x2 2
run cfg2_x4 4
run cfg2_x8 8
run cfg4_x8 8 --workload cfg4
run cfg5_x8 8 --workload cfg5
run cfg2_x8_queries 8 --parallelism queries
COBSGPU_OCC=2 COBSGPU_STAGES=12 run cfg2_x8_occ2 8
run cfg2_x8_nooverlap 8 --no-overlap
