# round 2: GPU tests, reference arm (whole batch), bench line (with the full-size parity check against the reference's
# cached index), ncu launch list of the same command, --set full captures of union_topk_kernel (the FULL C2 batch) and
# decode_score_kernel, the other workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_ref.json", "gpurun_out/bench_c2_final.json"):
    try:
        d = json.loads(open(f).read())
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d.get("parity_vs_reference"), (d.get("cpu_baseline") or {}).get("sample", "")[:160])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:union_topk -c 1 -o gpurun_out/prof_union_fullset python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_union_fullset.log 2>&1; echo "ncu union full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_score -c 1 -o gpurun_out/prof_decode_fullset python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_decode_fullset.log 2>&1; echo "ncu decode full rc=$?"
bash tools/workloads.sh
