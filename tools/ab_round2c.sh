python -m pytest tests/test_gpu_farneback.py tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -3
B="python bench.py --steps 40 --warmup 5 --no-cpu --no-parity"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d['value'], d['ms_per_step'], d['stage_ms_per_step'])" "$1"; }
$B 2>/dev/null | pick h512
DATMO_PYR_H_THREADS=256 $B 2>/dev/null | pick h256
DATMO_PYR_H_THREADS=384 $B 2>/dev/null | pick h384
DATMO_PYR_H_COLS=1 $B 2>/dev/null | pick h_cols
DATMO_DBSCAN_SUBTAGS=1 python tools/dbscan_prof.py 32
