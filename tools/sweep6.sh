mkdir -p gpurun_out
timeout 60 python tools/debug1.py "OR body 0 t0000001 t0000005 t0000020 t0000100 t0000500 t0001000" || exit 1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v7.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -12 gpurun_out/pytest_v7.log
[ $rc -ne 0 ] && exit 1
for cfg in "4 16" "4 20" "4 12"; do
  set -- $cfg
  timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --warps $1 --warps-per-sm $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG $1 $2', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'], d['roofline'].get('window_docs'), d['roofline'].get('doc_range_splits'))" || { echo "CFG $cfg FAILED"; exit 1; }
done
