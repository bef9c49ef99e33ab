mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v6.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -12 gpurun_out/pytest_v6.log
[ $rc -ne 0 ] && exit 1
for cfg in "4 16 0" "4 16 1" "4 20 1" "8 24 1"; do
  set -- $cfg
  timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --warps $1 --warps-per-sm $2 --generic-accumulate $3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG $1 $2 $3', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'], d['roofline'].get('window_docs'), d['roofline'].get('doc_range_splits'))" || { echo "CFG $cfg FAILED"; exit 1; }
done
