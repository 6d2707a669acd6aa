# full GPU validation: pytest -m gpu, smoke, final ncu captures of the two sweep kernels and the launch list of the bench command.
# usage: gpurun --timeout 1500 -- bash tools/gpu_final_validation.sh
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/profile_sweep.py c3 151552 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gsweep_kernel -s 1 -c 1 -f -o gpurun_out/r02_gsweep_c3 python tools/profile_sweep.py c3 151552 > gpurun_out/ncu_c3.log 2>&1
python tools/profile_gsweep.py 284160 10000 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gsweep_kernel -s 1 -c 1 -f -o gpurun_out/r02_gsweep_c5 python tools/profile_gsweep.py 284160 10000 > gpurun_out/ncu_c5.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out/r02_gsweep_c*.ncu-rep gpurun_out/r02_launches_c3.csv
