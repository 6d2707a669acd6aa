mkdir -p gpurun_out
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_try1.json 2> gpurun_out/bench_r02_try1.err; echo rc=$?
tail -c 600 gpurun_out/bench_r02_try1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_try1.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
for k,v in d.get('other_workloads',{}).items():
    print(k, json.dumps(v)[:700])
PY
