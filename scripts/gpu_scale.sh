#!/bin/bash
# scaling pass on one multi-GPU box: bench.py at the given rank counts (torchrun, NCCL)
mkdir -p gpurun_out
WL=${WL:-cfg2}
for N in "$@"; do
  for wl in $WL; do
    echo "== bench $wl x$N"
    if [ "$N" = "1" ]; then
      timeout -k 10 600 python bench.py --gpus 1 --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${wl}_x$N.json 2> gpurun_out/scale_${wl}_x$N.err
    else
      timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
        bench.py --gpus $N --workload $wl --steps 20 --warmup 3 > gpurun_out/scale_${wl}_x$N.json 2> gpurun_out/scale_${wl}_x$N.err
    fi
    echo "rc=$?"
    grep '^{' gpurun_out/scale_${wl}_x$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d value %.4g ms/step %.3f e2e %.4g K2 %.3f ms frac %.3f launches %d clocks %s'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['kernel_ms'],r['frac'],d['gpu_launches'],d['clocks']))"
    grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/scale_${wl}_x$N.err | tail -n 3
  done
done
