# parity of the union kernel paths + C2 bench lines for a few tunings (no ncu)
mkdir -p gpurun_out
DGPU_LIB=$PWD/diagon_b200/libdiagon_b200_chk.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "union or lane_merge3 or named or wide" > gpurun_out/pytest_union_chk.log 2>&1; rc=$?
echo "pytest(chk) rc=$rc"; tail -4 gpurun_out/pytest_union_chk.log
[ $rc -ne 0 ] && { grep -E "Error|error|assert" gpurun_out/pytest_union_chk.log | head -20; exit 1; }
bash tools/sweep.sh "$@"
