mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_step.py tests/test_gpu_testbed.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r3_pytest6.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3_pytest6.log
tail -3 gpurun_out/r3_pytest6.log
export NEUS2_BENCH_DATA_CACHE=/tmp/bench_data.npz
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r3_bench_$name.json 2> gpurun_out/r3_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3_bench_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],4), d.get('stages') or d.get('stage_ms'))
except Exception as e: print('$name', 'ERR', e)
PY
}
run bs3 A=1
run legacy3 NEUS2_LOSS_BY_SAMPLES=0
run bs4 A=1
