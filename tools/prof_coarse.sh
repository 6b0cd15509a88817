# SpeedOfLight / stall tables of the coarse-layer kernels of one step (run under gpurun)
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity"
$CMD1 > gpurun_out/pc_plain.log 2>&1 && \
ncu --section SpeedOfLight --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none \
    -k regex:"k_pyr_h|k_pyr_v|k_pyr0_polyexp_t|k_upsample|k_flow_iter_xm" -s 46 -c 23 -o gpurun_out/r02_prof_coarse $CMD1 > gpurun_out/pc_ncu.log 2>&1
tail -3 gpurun_out/pc_ncu.log
