#!/bin/bash
# End-of-round verification + profile refresh (run on the GPU box through gpurun): full GPU suite, smoke, the bench line, the
# launch list of the same command and one ncu --set full capture of the kernels that changed last (each after a clean run).
R=r02
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${R}_final_pytest.log 2>&1; echo "pytest rc $?" >> $O/${R}_final_pytest.log; tail -3 $O/${R}_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_final_smoke.log 2>&1; echo "smoke rc $?"
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err || exit 1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-variants > $O/${R}_bench_s1.json 2>/dev/null || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-variants > $O/${R}_ncu_launch.log 2>&1
cap() {  # name regex skip
  timeout 240 ncu --set full --clock-control none --kernel-name regex:$2 --launch-skip $3 --launch-count 1 --csv --page raw \
      --log-file $O/${R}_full_$1.csv python scripts/trace_step.py 16 mixed > $O/${R}_full_$1.log 2>&1
}
cap warp_assemble warp_assemble 41
cap wmedian wmedian 41
python scripts/trace_step.py 16 mixed > $O/${R}_trace.log 2>&1
ls -la $O | grep ${R}_ | tail -20
