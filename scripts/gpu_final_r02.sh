#!/bin/bash
# Final round-2 validation on ONE B200 (run under gpurun): full GPU suite, the default bench
# line of both arms, sanitizer checks of the pipelined exhaustive path, CLI wall time.
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r02_w_tests.txt
timeout 420 python bench.py > $O/r02_w_x1.json 2> $O/r02_w_x1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_w_ref.json 2> $O/r02_w_ref.err
timeout 150 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -x -q -k "pipelined and 5000" > $O/r02_w_memcheck.log 2>&1
timeout 150 compute-sanitizer --tool racecheck --kernel-regex-exclude kns=score_kernel python -m pytest tests/test_gpu_round2.py -x -q -k "pipelined and 5000" > $O/r02_w_racecheck.log 2>&1
GB=12 NQ=300000 timeout 200 python scripts/bench_cli.py > $O/r02_w_cli.txt 2>&1
tail -3 $O/r02_w_tests.txt $O/r02_w_memcheck.log $O/r02_w_racecheck.log $O/r02_w_cli.txt
