#!/bin/bash
# runs on the GPU box: bench every exp_*.so variant
for lib in qcrypto-ldpc_b200/exp_*.so; do
  n=$(basename $lib .so)
  QLDPC_LIB=$PWD/$lib python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e > gpurun_out/$n.json 2> gpurun_out/$n.err || tail -3 gpurun_out/$n.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/$n.json')); print('$n', round(d['value']), 'Mbit/s', round(d['ms_per_step'],2), 'ms')
except Exception as e: print('$n failed', e)"
done
