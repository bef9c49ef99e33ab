mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?
echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_gpu.log
bash tools/sweep.sh "C4|" "C2|"
timeout 900 ncu --section SourceCounters --section SchedulerStats --section WarpStateStats --section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --cache-control none --import-source on -k regex:union_topk -c 1 -o gpurun_out/prof_union_c4 python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_union_c4.log 2>&1; echo "ncu rc=$?"
