#!/bin/bash
# Round profile captures, part 2: one `ncu --set full` launch per kernel, taken at the last (full-resolution, B = 16)
# launch of that kernel inside one classic+nl-fast step.  Raw metric pages go to CSV (small); only the dominant kernel
# also keeps its .ncu-rep (source page).
R=${1:-r01}
O=gpurun_out
python scripts/trace_step.py 16 mixed > /dev/null 2>&1 || exit 1
cap() {  # name regex skip
  timeout 240 ncu --set full --clock-control none --kernel-name regex:$2 --launch-skip $3 --launch-count 1 --csv --page raw \
      --log-file $O/${R}_full_$1.csv python scripts/trace_step.py 16 mixed > $O/${R}_full_$1.log 2>&1
}
cap warp_assemble warp_assemble 41
cap wmedian wmedian 41
cap occlusion occlusion 41
cap rof_iter rof_iter 300
cap level_prep level_prep 13
cap gauss_resize gauss_resize 20
cap clip_add clip_add 41
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name pcg_ic_kernel --launch-skip 41 --launch-count 1 \
    -o $O/${R}_pcg_ic -f python scripts/trace_step.py 16 mixed > $O/${R}_full_pcg.log 2>&1
ls -la $O | grep ${R}_ | tail -20
