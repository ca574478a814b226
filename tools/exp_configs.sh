#!/bin/bash
python -m pytest tests/test_reference_gpu.py -m gpu -q -x -k "full_position_ids_describe" 2>&1 | tail -15
for c in C1 C3; do python bench.py --config $c --no-cpu-baseline --no-extras 2>/dev/null | grep "^{" > gpurun_out/bench_r02_$c.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_$c.json')); print('$c', d['ms_per_step'], d['e2e']['ttft_ms'], d['decode_ms_per_token'], d['virtual_splice'], d['ms_per_step_with_copied_splice'], d['roofline']['frac'])"; done
