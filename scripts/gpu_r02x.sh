#!/bin/bash
# r02x visit (1 GPU): 160-thread nearest-neighbour CTAs as the default for 129..160 and 257..320 columns — GPU tests,
# smoke, parity sweep, and the BASELINE shapes against the 256-thread form
TAG=r02x
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 > $O/${TAG}_pytest.txt; tail -5 $O/${TAG}_pytest.txt | cut -c1-600
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/${TAG}_smoke.txt
for nt in 0 256; do
  echo "== ACB200_NN_NT=$nt (0 = default widths)"
  ACB200_NN_NT=$nt MEASURE_SCALES=nn timeout 300 python scripts/measure_configs.py 2>&1 | python -c "
import sys, ast
for l in sys.stdin:
    if l.startswith('{'):
        d = ast.literal_eval(l); print('   %-45s %-5s %.4f ms per %d frames' % (d['config'], d['content'], d['ms_per_pass'], d['frames']))
"
done | tee $O/${TAG}_nn_nt.txt
echo "== fuzz 45 s"; timeout 300 python scripts/fuzz_parity.py 45 2>&1 | tail -3 | tee $O/${TAG}_fuzz_parity.txt
