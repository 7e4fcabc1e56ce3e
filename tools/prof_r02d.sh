# Round-2, last pass: the lighter capture kernel (4 warps x 3 stages, per-lane masks).  Every ncu run follows a plain run of
# the same command that exited 0.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-graphs --no-configs --no-cpu-baseline --no-kernels"
$CMD > gpurun_out/r02d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/r02d_launches.csv $CMD > gpurun_out/r02d_ncu_l.log 2>&1
$CMD > gpurun_out/r02d_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_x_capture_tma' -s 4 -c 3 -o gpurun_out/r02d_prof_capture $CMD > gpurun_out/r02d_ncu_f.log 2>&1
ls -la gpurun_out/r02d_prof_capture.ncu-rep
python bench.py > gpurun_out/r02d_bench_1gpu.jsonl 2> gpurun_out/r02d_bench_1gpu.err
tail -c 300 gpurun_out/r02d_bench_1gpu.err
