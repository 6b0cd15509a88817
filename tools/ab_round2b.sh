# A/B of the round-2 coarse-layer kernels against the ones they replace (run under gpurun)
python -m pytest tests/test_gpu_farneback.py -x -q -m gpu > gpurun_out/ab_tests.log 2>&1; tail -4 gpurun_out/ab_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu --no-parity"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d['value'], d['ms_per_step'], d['stage_ms_per_step'])" "$1"; }
$B 2>/dev/null | pick new
DATMO_PYR_H_COLS=1 $B 2>/dev/null | pick pyr_h_cols
DATMO_PYR_V_OLD=1 $B 2>/dev/null | pick pyr_v_old
DATMO_POLYEXP_GENERIC=1 $B 2>/dev/null | pick polyexp_generic
DATMO_XM_TY=46 $B 2>/dev/null | pick xm_ty46
DATMO_XM_TY=46 DATMO_PYR_H_COLS=1 DATMO_PYR_V_OLD=1 DATMO_POLYEXP_GENERIC=1 $B 2>/dev/null | pick all_old
python tools/configs_bench.py 2>&1 | tail -6
