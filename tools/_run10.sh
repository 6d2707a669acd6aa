mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; echo rc=$?
tail -c 400 gpurun_out/bench_r02_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'n',d['n_gpus'])
print(json.dumps(d.get('other_workloads',{}))[:3000])
PY
