# Round-2, second pass: the TMA-fed background pass.  Every ncu run follows a plain run of the same command that exited 0.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-graphs --no-configs --no-cpu-baseline --no-kernels"
$CMD > gpurun_out/r02b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_l.log 2>&1
$CMD > gpurun_out/r02b_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_x_capture_tma|k_xw_ell' -s 4 -c 4 -o gpurun_out/r02b_prof_capture $CMD > gpurun_out/r02b_ncu_f.log 2>&1
ls -la gpurun_out/r02b_prof_capture.ncu-rep
python tools/timeline.py > gpurun_out/r02b_timeline_step.txt 2>&1
tail -4 gpurun_out/r02b_timeline_step.txt
