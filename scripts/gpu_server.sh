#!/bin/bash
# GPU visit for the resident-source server path: its tests, then timings beside the compiled reference
O=gpurun_out; mkdir -p $O
echo "== pytest mixed"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "mixed or composite" 2>&1 | tail -15 | cut -c1-300
echo "== diag_server"; timeout 300 python scripts/diag_server.py 2>&1 | tail -8
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -5 | cut -c1-300
