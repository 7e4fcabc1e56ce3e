# Round-2 closing pass at HEAD: launch list of the bench command (after a plain run that exited 0), step timelines, the full bench line.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-graphs --no-configs --no-cpu-baseline --no-kernels"
$CMD > gpurun_out/r02e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/r02e_launches.csv $CMD > gpurun_out/r02e_ncu_l.log 2>&1
python tools/timeline.py > gpurun_out/r02e_timeline_step.txt 2>&1
python tools/timeline_pheme.py 24 > gpurun_out/r02e_timeline_pheme24.txt 2>&1
python tools/timeline_pheme.py 4096 > gpurun_out/r02e_timeline_pheme4096.txt 2>&1
python bench.py > gpurun_out/r02e_bench_1gpu.jsonl 2> gpurun_out/r02e_bench_1gpu.err
tail -c 200 gpurun_out/r02e_bench_1gpu.err
