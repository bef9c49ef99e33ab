# C4 (OR-5 + range filter, top-100): bench line and a --set full capture of the mode-2 instantiation at full size
mkdir -p gpurun_out
timeout 300 python bench.py --workload C4 --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/wl_C4.err | tee gpurun_out/wl_C4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:union_topk -c 1 -o gpurun_out/prof_c4_final python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4_final.log 2>&1; echo "ncu rc=$?"
