# A/B of the two disjunction kernels on the bench workload (exits on the first failure)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && exit 1
for lm in 1 0; do
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --lane-merge $lm 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('LM$lm', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])" || exit 1
done
