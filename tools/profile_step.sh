#!/bin/bash
# Round profile of the BASELINE train step on one B200 (run through gpurun from the repo root):
#   full bench line, reference arm, ncu launch list of 4 steps, ncu --set full of one step's own kernels -> compact table.
# 63 launches of a train step match the -k filter below (count them in the launch list when kernels are added): skipping 3 x 63 captures
# exactly the fourth step, forward to Adam.  (The round-1 capture used --launch-skip 210 -c 70 and therefore starts at the decoder forward.)
# The .ncu-rep stays on the box (too large for gpurun_out); tools/ncu_table.py prints the table that is committed under profiles/.
set -u
TAG=${1:-v2}
python bench.py > gpurun_out/r1_bench_n1_$TAG.json 2> gpurun_out/r1_bench_n1_$TAG.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1_bench_ref_$TAG.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1_launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-eval --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none \
    -k regex:"fold_|bn_leaky|bn_bwd2|norm_stats|gemm_rows|gemm_wgrad|gemm_vn|nn_prefilter|maxpool_argmax|rows_sample_sum|pool_linear|nn_exact|nn_resolve" \
    --launch-skip 189 -c 63 -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-eval --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
python tools/ncu_table.py /tmp/prof_$TAG.ncu-rep > gpurun_out/r1_ncu_step_$TAG.txt
ls -la /tmp/prof_$TAG.ncu-rep
echo done
