#!/bin/bash
# r02w visit (1 GPU): nearest-neighbour kernel with 160 / 320 threads per CTA for 320-column rows vs the 256-thread default
TAG=r02w
O=gpurun_out
mkdir -p $O
for nt in 256 320 160; do
  echo "== ACB200_NN_NT=$nt"
  ACB200_NN_NT=$nt MEASURE_ONLY=320x96 MEASURE_SCALES=nn timeout 300 python scripts/measure_configs.py 2>&1 | python -c "
import sys, ast
for l in sys.stdin:
    if l.startswith('{'):
        d = ast.literal_eval(l); print('   %-45s %-5s %.4f ms per 256 frames' % (d['config'], d['content'], d['ms_per_pass']))
"
done | tee $O/${TAG}_nn_nt.txt
echo "== parity with 320 and 160 threads"; for nt in 320 160; do ACB200_NN_NT=$nt timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x --tb=short 2>&1 | tail -2; done | tee $O/${TAG}_pytest_nn_nt.txt
