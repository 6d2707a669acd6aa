mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/r02_bench_default_1gpu.err; echo rc=$?
tail -c 300 gpurun_out/r02_bench_default_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default_1gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], 'cpu', d.get('cpu_baseline'))
for k,v in d.get('other_workloads',{}).items():
    print(k, json.dumps(v)[:420])
print(open('gpurun_out/r02_bench_reference_arm.json').read()[:600])
PY
