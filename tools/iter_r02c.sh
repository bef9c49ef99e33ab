mkdir -p gpurun_out
one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])"; }
timeout 300 python bench.py --workload C5 --scale 0.1 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | one "C5 now "
for w in 16384 32768 65536; do
timeout 300 python bench.py --workload C4 --union-window-docs $w --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | one "C4 W=$w "
done
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | one "C2 now "
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search_after or tuning or option" 2>&1 | tail -2
