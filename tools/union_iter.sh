# iteration loop for union_topk_kernel: parity (bounds-checked build), bench lines, per-line counters at full size
mkdir -p gpurun_out
DGPU_LIB=$PWD/diagon_b200/libdiagon_b200_chk.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "union or lane_merge3 or named or wide" > gpurun_out/pytest_union_chk.log 2>&1; rc=$?
echo "pytest(chk) rc=$rc"; tail -4 gpurun_out/pytest_union_chk.log
[ $rc -ne 0 ] && { grep -E "Error|error|assert" gpurun_out/pytest_union_chk.log | head -20; exit 1; }
for cfg in "--lane-merge 3" "--lane-merge 3 --union-window-docs 65536" "--lane-merge 3 --union-window-docs 16384" ${EXTRA_CFGS}; do
  name=$(echo $cfg | tr -d ' -' )
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $cfg > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$cfg" gpurun_out/bench_$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    print('CFG', sys.argv[1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline'].get('step_ms_by_kernel'), 'frac', d['roofline'].get('frac'))
except Exception as e:
    print('CFG', sys.argv[1], 'FAILED', e)
PY
done
bash tools/union_ncu_full.sh
bash tools/workloads.sh
