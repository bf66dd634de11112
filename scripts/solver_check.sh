#!/bin/bash
# solver iteration loop: parity tests of the solver / pipeline, then the bench step
python -m pytest tests/test_gpu_solver.py tests/test_gpu_rowband.py tests/test_gpu_pipeline.py -q -x 2>&1 | tail -2
python -m pytest tests/test_gpu_config_goldens.py -q -x -k "middlebury and (Grove3 or Urban3) or bench_workload" 2>&1 | tail -1
python bench.py --steps 3 --no-cpu-baseline --no-configs --no-variants 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pairs/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'solver ms', round(d['stages']['solver']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4), 'iters', d['pcg_iters_per_step'], 'nc', d['pcg_not_converged'])"
