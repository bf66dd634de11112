#!/bin/bash
# Experiment (round 2): concurrent sub-batches -- does the HBM-bound solver of one group overlap the issue-bound weighted
# median of the other?  Run on the GPU box through gpurun; writes gpurun_out/exp_split_*.json
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B --split 1 > $O/exp_split_1.json 2> $O/exp_split_1.err
$B --split 2 > $O/exp_split_2.json 2> $O/exp_split_2.err
B200FLOW_NO_SOLVER_PRIORITY=1 $B --split 2 > $O/exp_split_2_noprio.json 2> $O/exp_split_2_noprio.err
B200FLOW_LIB=$PWD/optical-flow-python_b200/variants/lib_wm4.so $B --split 2 > $O/exp_split_2_wm4.json 2> $O/exp_split_2_wm4.err
$B --split 2 --batch 32 > $O/exp_split_2_b32.json 2> $O/exp_split_2_b32.err
for f in $O/exp_split_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","stage_ms_per_step","pcg_iters_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["measured"])
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
