for c in 32 64 128 1024; do
  timeout 100 python bench.py --workload C3-AND4 --steps 5 --warmup 3 --no-cpu-baseline --decode-ctas-per-sm $c 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG C3 dec_ctas=$c', round(d['value']), d['roofline']['step_ms_by_kernel']['decode_score_kernel'], round(d['roofline']['decode_score_kernel']['achieved']))"
done
for c in 32 128; do
  timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --decode-ctas-per-sm $c 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG C2 dec_ctas=$c', round(d['value']), d['roofline']['step_ms_by_kernel']['decode_score_kernel'], round(d['roofline']['decode_score_kernel']['achieved']))"
done
