#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun).  Every ncu pass runs only after the same command has exited 0
# without ncu; numbers printed under ncu are never bench values.
R=r02
O=gpurun_out
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference_arm.json 2> $O/${R}_ref.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-variants > $O/${R}_bench_s1.json 2>/dev/null || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-variants > $O/${R}_ncu_launch.log 2>&1
python scripts/trace_step.py 16 mixed > $O/${R}_trace.log 2>&1 || exit 1
cap() {  # name regex skip
  timeout 240 ncu --set full --clock-control none --kernel-name regex:$2 --launch-skip $3 --launch-count 1 --csv --page raw \
      --log-file $O/${R}_full_$1.csv python scripts/trace_step.py 16 mixed > $O/${R}_full_$1.log 2>&1
}
cap warp_assemble warp_assemble 41
cap wmedian wmedian 41
cap occlusion occlusion 41
cap rof_tile rof_tile 40
cap level_prep level_prep 13
cap gauss_resize gauss_resize 20
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name pcg_ic_kernel --launch-skip 41 --launch-count 1 \
    -o $O/${R}_pcg_ic -f python scripts/trace_step.py 16 mixed > $O/${R}_full_pcg.log 2>&1
ncu -i $O/${R}_pcg_ic.ncu-rep --page raw --csv > $O/${R}_full_pcg_ic.csv 2>/dev/null
# local-memory traffic of the solver (the SASS carries a few STL / LDL at the 128-register cap): how many are executed?
timeout 240 ncu --metrics smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed.sum,gpu__time_duration.sum \
    --clock-control none --kernel-name pcg_ic_kernel --launch-skip 41 --launch-count 1 --csv --log-file $O/${R}_pcg_ic_local.csv \
    python scripts/trace_step.py 16 mixed > /dev/null 2>&1
python scripts/ncu_summary.py $O/${R}_full_*.csv > $O/${R}_ncu_full_kernels.jsonl 2>&1
python scripts/pcg_bench.py --solver 4 > $O/${R}_pcgbench_ic.jsonl 2>&1
ls -la $O | grep ${R}_ | tail -30
