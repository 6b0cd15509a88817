#!/bin/bash
# Round-2 measurement protocol (run under gpurun from the repo root; writes gpurun_out/r02_*).
# Every ncu command follows a plain run of the same command line that exited 0.
set -u
O=gpurun_out
mkdir -p $O
python __graft_entry__.py --smoke > $O/r02_smoke.log 2>&1; tail -1 $O/r02_smoke.log
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; tail -c 300 $O/r02_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_n1.json 2> $O/r02_bench_reference_n1.err
python tools/configs_bench.py > $O/r02_configs.txt 2>&1
python tools/latency.py > $O/r02_latency.txt 2>&1
python tools/preprocess_latency.py > $O/r02_preprocess.txt 2>&1
DATMO_DBSCAN_SUBTAGS=1 python tools/dbscan_prof.py 32 > $O/r02_dbscan_stage_table.txt 2>&1
python bench.py --workload cfg5 --steps 20 --warmup 3 > $O/r02_cfg5_n1.json 2> $O/r02_cfg5_n1.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity"
$CMD > $O/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches.csv $CMD > $O/r02_ncu_ll.log 2>&1
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity"
$CMD1 > $O/r02_plain1.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:k_flow_iter_xm -s 40 -c 1 -o $O/r02_prof_xm $CMD1 > $O/r02_ncu_xm.log 2>&1
$CMD1 > $O/r02_plain1.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:k_pyr0_polyexp_t -s 5 -c 1 -o $O/r02_prof_pyr0 $CMD1 > $O/r02_ncu_pyr0.log 2>&1
$CMD1 > $O/r02_plain1.log 2>&1 && \
ncu --section SpeedOfLight --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis --clock-control none \
    -k regex:"k_run_|k_velmask|k_cluster|k_chain|k_pyr_h|k_pyr_v|k_pyr0_polyexp_t|k_upsample" -s 69 -c 46 -o $O/r02_prof_small $CMD1 > $O/r02_ncu_small.log 2>&1
PRE="python tools/preprocess_latency.py"
$PRE > $O/r02_plain2.log 2>&1 && \
ncu --section SpeedOfLight --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis --clock-control none \
    -k regex:"k_pre_accum|k_ransac_bounds|k_ransac_score|k_bev" -s 60 -c 8 -o $O/r02_prof_pre $PRE > $O/r02_ncu_pre.log 2>&1
ls -la $O/r02_* | awk '{print $5, $9}'
