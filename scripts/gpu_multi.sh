#!/bin/bash
# multi-GPU pass: N ranks under torchrun, document-sharded cfg2 (and cfg4), NCCL all-gather + merge
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
for wl in cfg2 cfg4; do
  echo "== bench $wl x$N"
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_x$N.json 2> gpurun_out/bench_${wl}_x$N.err
  echo "rc=$?"
  grep '^{' gpurun_out/bench_${wl}_x$N.json; tail -n 5 gpurun_out/bench_${wl}_x$N.err
done
echo "== reference arm under torchrun"
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_x$N.json 2> gpurun_out/bench_ref_x$N.err
echo "rc=$?"; grep '^{' gpurun_out/bench_ref_x$N.json | cut -c1-300
echo "== sharded parity on N gpus (cobs query --gpus N vs --gpus 1)"
G=tests/golden
./build/cobs query -i $G/random203.cobs_classic -t 0.02 --gpus 1 $(python -c "from oracle import oracle;print(oracle.random_query(5,200).decode())") > gpurun_out/cli_g1.txt 2>/dev/null
./build/cobs query -i $G/random203.cobs_classic -t 0.02 --gpus $N $(python -c "from oracle import oracle;print(oracle.random_query(5,200).decode())") > gpurun_out/cli_gN.txt 2>/dev/null
cmp gpurun_out/cli_g1.txt gpurun_out/cli_gN.txt && echo "cli sharded == unsharded ($(wc -l < gpurun_out/cli_g1.txt) lines)"
