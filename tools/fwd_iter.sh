#!/bin/bash
# one GPU iteration on the tensor-core forward: its parity tests, kernel timings, and the shared-memory counters of one launch
python -m pytest tests/test_gpu_parity.py -q -x -k "tc or tensor" 2>&1 | tail -3
python tools/time_kernels.py timit_c2 2>&1 | tail -1
ncu --clock-control none -k regex:eodm_tc_fwd3_kernel -s 3 -c 1 --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max python tools/time_kernels.py timit_c2 2>&1 | grep -E "^\s+(gpu__|l1tex|smsp|sm__)"
