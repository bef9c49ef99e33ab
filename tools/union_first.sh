# first contact of union_topk_kernel with a GPU: parity with device-side bounds checks, then A/B bench lines
mkdir -p gpurun_out
DGPU_LIB=$PWD/diagon_b200/libdiagon_b200_chk.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "union or lane_merge=3" > gpurun_out/pytest_union_chk.log 2>&1; rc=$?
echo "pytest(chk) rc=$rc"; tail -15 gpurun_out/pytest_union_chk.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "union or lane_merge=3 or named" > gpurun_out/pytest_union.log 2>&1; rc2=$?
echo "pytest rc=$rc2"; tail -5 gpurun_out/pytest_union.log
for cfg in "--lane-merge 1" "--lane-merge 3" "--lane-merge 3 --union-window-docs 65536" "--lane-merge 3 --union-window-docs 16384" "--lane-merge 3 --union-window-docs 49152"; do
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
