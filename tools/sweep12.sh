for c in 1 2 4 8 16; do
  timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --part-factor $c 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG part_factor=$c', round(d['value']), d['roofline']['step_ms_by_kernel'], d['roofline']['work_items'])"
done
