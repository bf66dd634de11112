#!/bin/bash
# weighted-median iteration loop: bit-exactness tests, then the filter-stage time of the bench step
python -m pytest tests/test_gpu_stages.py -q -x -k "median or wmed or weighted" 2>&1 | tail -2
python -m pytest tests/test_gpu_config_goldens.py -q -x -k "middlebury and (Grove3 or Urban3 or RubberWhale)" 2>&1 | tail -1
python bench.py --steps 3 --no-cpu-baseline --no-configs --no-variants 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pairs/s', round(d['value'],2), 'ms/step', round(d['ms_per_step'],1), 'wmedian ms', round(d['stages']['weighted_median']['ms_per_step'],2))"
