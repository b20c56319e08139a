#!/bin/bash
# Round-2 profile of the BASELINE train step on one B200 (run through gpurun from the repo root): ncu --set full of ONE train step's own
# kernels (58 launches of a step match the -k filter: skipping 3 x 58 captures exactly the fourth step) -> compact table under gpurun_out/.
# The .ncu-rep stays on the box (too large for gpurun_out); tools/ncu_table.py prints the table that is committed under profiles/.
set -u
FILTER='regex:gemm_rows|gemm_wgrad|gemm_vn|tail_dgrad|bwd1_p2|bn_bwd2|nn_prefilter|nn_resolve|nn_exact|fold_|bn_leaky_dot|bn_leaky_fwd|maxpool_argmax|pool_linear|rows_sample'
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-eval --no-cpu-baseline --no-chamfer-leg"
$CMD > gpurun_out/plain_f.log 2>&1 && \
ncu --set full --clock-control none -k "$FILTER" --launch-skip 174 -c 58 -o /tmp/prof_r2 $CMD > gpurun_out/ncu_f.log 2>&1
python tools/ncu_table.py /tmp/prof_r2.ncu-rep > gpurun_out/r2_ncu_step.txt
ls -la /tmp/prof_r2.ncu-rep
echo done
