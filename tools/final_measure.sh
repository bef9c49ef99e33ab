# the round's final numbers: GPU tests, reference arm, bench line, ncu launch list of the same command, and one
# --set full capture of the dominant kernel on a scaled C2 (the full-size launch does not finish under ncu replay)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1; echo "ncu rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:staged_merge -c 1 -o gpurun_out/prof_final python bench.py --scale 0.1 --queries 1000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1; echo "ncu full rc=$?"
