mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -x -q > gpurun_out/r3_pytest_dp.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3_pytest_dp.log
tail -3 gpurun_out/r3_pytest_dp.log
export NEUS2_BENCH_DATA_CACHE=/tmp/bench_data.npz
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r02_bench_n2_final2.json 2> gpurun_out/r02_bench_n2_final2.err; echo "bench2 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n2_final2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('stage_ms'))"
