#!/bin/bash
# Round-2 evidence run (one gpurun call): reference arm from the staged archive, bench, ncu launch list of the same
# command, one full capture of the headline kernels, the large-D VJP kernels, the large-D timing sweep.
set -x
O=gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_reference.json 2> $O/r2f_bench_reference.err
python bench.py > $O/r2f_bench.json 2> $O/r2f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2f_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check > $O/r2f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rk4_fwd_h_kernel|rk4_bwd_mma_kernel|param_grad_kernel|whiten' \
    -s 6 -c 5 -o $O/r2f_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-graph > $O/r2f_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rff_vjp_large_kernel|vjp_large_kernel' -s 2 -c 2 \
    -o $O/r2f_large64 python tools/large_vjp_once.py 64 > $O/r2f_ncu_large.log 2>&1
python tools/time_large_bwd.py --out $O/r2f_large_d.json > $O/r2f_large_d.log 2>&1
tail -2 $O/r2f_large_d.log | cut -c1-300
