mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v3b.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_v3b.log
[ $rc -ne 0 ] && exit 1
for cfg in "3 4 2048" "6 4 2048" "4 4 2048" "3 8 2048" "3 4 4096" "4 8 1024"; do
  set -- $cfg
  timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --queries 2000 --ctas-per-sm $1 --warps $2 --ring-entries $3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG $1 $2 $3', round(d['value']), d['roofline']['step_ms_by_kernel'], d['roofline'].get('window_docs'))" || { echo "CFG $cfg FAILED"; exit 1; }
done
