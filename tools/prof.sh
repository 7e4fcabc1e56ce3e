set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernels"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 260 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_xw_scan|k_prop1_mix|k_dw_sweep|k_bwd_mix|k_propagate$' -s 5 -c 6 -o gpurun_out/prof_r01b $CMD > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/prof_r01b.ncu-rep; tail -3 gpurun_out/ncu_f.log | cut -c1-300
