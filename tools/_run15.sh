mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo gpus=$N
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo rc=$?
tail -c 300 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'n',d['n_gpus'])
c=d['other_workloads']['c5_g_sweep_1e7_runs_strong_scaled']
print('C5 linear',c['linear']['value'],c['linear']['ms_per_sweep'],'e2e',c['linear']['e2e']['value'],'g',c['linear']['g_T']); print(c['result_check'])
PY
