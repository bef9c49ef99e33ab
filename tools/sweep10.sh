for cfg in "20 5" "20 4" "20 3" "24 3" "24 4" "28 3"; do
  set -- $cfg
  timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --warps-per-sm $1 --stage-log2 $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG wps=$1 stage=$2', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel']['accumulate_topk_kernel+intersect_topk_kernel'], d['roofline'].get('window_docs'))" || { echo "CFG $cfg FAILED"; }
done
