#!/bin/bash
# one GPU iteration on the tensor-core VJP / tail: parity tests, role profile, brief bench
python -m pytest tests/test_gpu_parity.py -q -x -k "tc or tensor or session or graph" 2>&1 | tail -2
python tools/tcb_profile.py timit_c2 2>&1 | head -3
bash tools/bench_brief.sh
