#!/bin/bash
# gpurun_out/r02_* (written by tools/collect_profiles.sh on the GPU box) -> the tracked files under profiles/
set -e
O=gpurun_out P=profiles
for f in bench_n1.json bench_reference_n1.json cfg5_n1.json configs.txt latency.txt preprocess.txt dbscan_stage_table.txt launches.csv; do cp $O/r02_$f $P/r02_$f; done
python tools/launches.py $P/r02_launches.csv > $P/r02_launches_summary.txt
python tools/ncu_summary.py $O/r02_prof_xm.ncu-rep > $P/r02_flow_iter_xm_ncu_full.txt
python tools/ncu_summary.py $O/r02_prof_pyr0.ncu-rep > $P/r02_pyr0_polyexp_ncu_full.txt
python tools/ncu_summary.py $O/r02_prof_small.ncu-rep > $P/r02_small_kernels_ncu.txt
python tools/ncu_summary.py $O/r02_prof_pre.ncu-rep > $P/r02_preprocess_kernels_ncu.txt
for k in xm:flow_iter_xm pyr0:pyr0_polyexp; do
  ncu -i $O/r02_prof_${k%%:*}.ncu-rep --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
  python tools/ncu_source.py /tmp/_src.csv > $P/r02_${k##*:}_source_hotspots.txt
done
ncu -i $O/r02_prof_xm.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); hdr,units,vals=rows[0],rows[1],rows[2]
def get(name):
    i=hdr.index(name); v=float(vals[i].replace(',',''))
    return v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9,'ns':1e-3,'us':1,'usecond':1,'ms':1e3}.get(units[i],1)
rd,wr,dur=get('dram__bytes_read.sum'),get('dram__bytes_write.sum'),get('gpu__time_duration.sum')
json.dump({'kernel':'void <unnamed>::k_flow_iter_xm<<unnamed>::XmTile<46, 320, 2, 2, 2>>','grid':'(2, 23, 32)','block':'(320, 1, 1)',
 'dram_bytes_per_launch':rd+wr,'dram_read_bytes':rd,'dram_write_bytes':wr,'duration_us_under_ncu':dur,
 'capture':'r02_prof_xm.ncu-rep (tools/collect_profiles.sh)',
 'launch':'level-0 iteration (1024x1024), 32 pairs: algorithmic bytes of this launch = 56 B x 1024 x 1024 x 32 = 1879048192'},
 open('profiles/roofline_traffic.json','w'),indent=1)"
bash tools/sass_excerpts.sh > $P/r02_sass_excerpts.txt
sed -i 's/k_run_link  /k_run_heads /; s/k_run_flatten (x2)  /k_run_flatten_list  /; s/k_run_union rows 2..r       /k_run_pairs (row 0, rows 1..r)/' $P/r02_dbscan_stage_table.txt
