#!/bin/bash
# Round profile captures, part 1 (run on the GPU box through gpurun): bench lines + ncu launch list of the bench command.
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
R=${1:-r01}
O=gpurun_out
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err || exit 1
python bench.py --solver-precision mixed-jacobi --no-cpu-baseline > $O/${R}_bench_jacobi.json 2> $O/${R}_bench_jacobi.err || exit 1
python bench.py --solver-precision fp64 --no-cpu-baseline > $O/${R}_bench_fp64.json 2> $O/${R}_bench_fp64.err || exit 1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${R}_bench_s1.json 2>/dev/null || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${R}_ncu_launch.log 2>&1
python scripts/trace_step.py 16 mixed > $O/${R}_trace.log 2>&1
python scripts/pcg_bench.py --solver 4 > $O/${R}_pcgbench_ic.log 2>&1
python scripts/pcg_bench.py --solver 0 > $O/${R}_pcgbench_jacobi.log 2>&1
ls -la $O | tail -8
