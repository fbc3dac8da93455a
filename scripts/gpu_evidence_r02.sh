#!/bin/bash
# Round-2 evidence on ONE B200 (run under gpurun): ncu launch list + full capture of the score
# kernel on the BENCHED 131 GB matrix, sanitizer race/sync checks, a sustained (>= 5 s) run.
# Outputs under gpurun_out/; summaries are copied into profiles/ afterwards.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --no-secondary --no-load --no-cpu-baseline --no-parity"

# 1. every launch of a short default-workload run with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/r02_launches_cfg4.csv $B --no-patterns --steps 2 --warmup 3 > $O/r02_launches_cfg4.json 2> $O/r02_launches_cfg4.err

# 2. the score kernel of the headline leg (CAND, 8 planes), full set, on the benched geometry
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 1 \
    -o $O/r02_score_cand_cfg4 $B --no-patterns --steps 2 --warmup 3 > /dev/null 2> $O/r02_ncu_cand.err
# 2b. the top-k epilogue variant and the 16-plane variant (patterns legs come after 2 x (3+2)+... launches)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:score_kernel<.int.3, .int.3, .int.8>' -c 1 \
    -o $O/r02_score_topk_cfg4 $B --steps 2 --warmup 3 > /dev/null 2> $O/r02_ncu_topk.err
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:score_kernel<.int.3, .int.3, .int.16>' -c 1 \
    -o $O/r02_score_topk16_cfg4 $B --steps 2 --warmup 3 > /dev/null 2> $O/r02_ncu_topk16.err

# 3. sustained run: 100 steps x 57 ms, clocks sampled throughout
$B --no-patterns --steps 100 --warmup 3 > $O/r02_sustained_cfg4.json 2> $O/r02_sustained_cfg4.err

# 4. race / sync checks of the mbarrier pipeline on small inputs
compute-sanitizer --tool racecheck --racecheck-report all python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_sanitizer_racecheck.log 2>&1
compute-sanitizer --tool racecheck --kernel-regex-exclude kns=score_kernel python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_sanitizer_racecheck_other_kernels.log 2>&1
compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_sanitizer_synccheck.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -x -q -k "topk_all_equal or submit_collect or exhaustive_lists_counting_sort and 5000" > $O/r02_sanitizer_memcheck.log 2>&1
tail -3 $O/r02_sanitizer_racecheck.log $O/r02_sanitizer_synccheck.log $O/r02_sanitizer_memcheck.log
ls -la $O/*.ncu-rep
