set -x
for cfg in "3 8" "6 4" "2 8" "1 8" "4 4"; do
  set -- $cfg
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --queries 2000 --ctas-per-sm $1 --warps $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG $1 $2', d['value'], d['roofline']['step_ms_by_kernel'], d['roofline'].get('window_docs'))"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:accumulate_topk -c 1 -o gpurun_out/prof_v3a python bench.py --steps 1 --warmup 1 --no-cpu-baseline --queries 500 > gpurun_out/ncu_v3a.log 2>&1
echo ncu rc=$?
