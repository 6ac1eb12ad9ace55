mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3_pytest2.log
tail -5 gpurun_out/r3_pytest2.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r3_bench_$name.json 2> gpurun_out/r3_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3_bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), d.get('stages') or d.get('stage_ms'))
except Exception as e: print('$name', 'ERR', e)
PY
}
L=/root/repo/neus2_b200/csrc
run t5s3 A=1
run t5s2 NEUS2_TAIL_GEN_SHARE=2
run t5s4 NEUS2_TAIL_GEN_SHARE=4
run t6s3 NEUS2_B200_LIB=$L/build_tail6/libneus2_b200.so
run t6s2 NEUS2_B200_LIB=$L/build_tail6/libneus2_b200.so NEUS2_TAIL_GEN_SHARE=2
run t6s4 NEUS2_B200_LIB=$L/build_tail6/libneus2_b200.so NEUS2_TAIL_GEN_SHARE=4
run t4s2 NEUS2_B200_LIB=$L/build_tail4/libneus2_b200.so
run sep NEUS2_TAIL_MERGED=0
