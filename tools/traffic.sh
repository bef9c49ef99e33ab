mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:accumulate_topk -c 3 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -E "dram__bytes_read|hit_rate|time_duration" gpurun_out/traffic.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | head -12
