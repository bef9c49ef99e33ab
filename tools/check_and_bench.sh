mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && exit 1
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'], d['roofline'].get('window_docs'), d['roofline'].get('doc_range_splits'))"
