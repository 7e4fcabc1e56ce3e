# ncu evidence for propagate / readout on a forest far larger than L2 (4.2 M rows, 1.07 GB per panel)
set -x
CMD="python tools/kbench.py --forest=uniform256"
$CMD > gpurun_out/kbench_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_propagate$|k_readout_part' -s 8 -c 3 -o gpurun_out/prof_r01b_kbench $CMD > gpurun_out/ncu_kb.log 2>&1
cat gpurun_out/kbench_plain.log | tail -6
# propagate: launches 0..13 are the parent-gather (TD) orientation, 14..27 the child segment-sum (BU)
ncu --set full --clock-control none --import-source on -k 'regex:k_propagate$' -s 12 -c 4 -o gpurun_out/prof_r01b_kbench_prop $CMD > gpurun_out/ncu_kb2.log 2>&1
