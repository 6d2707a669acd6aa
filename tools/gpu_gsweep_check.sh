# gsweep kernels on one B200: parity against the oracle (d = 1 bit-exact, d = 5 single-pass), timing vs the round-1 kernel, ncu instruction counts.
# usage: gpurun -- bash tools/gpu_gsweep_check.sh
mkdir -p gpurun_out
python tools/try_c3.py parity 1000000 > gpurun_out/g_c3.log 2>&1
python tools/try_gsweep.py parity time >> gpurun_out/g_c5.log 2>&1
python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/g_tests.log
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:gsweep_kernel -s 1 -c 1 --csv --log-file gpurun_out/g_c3_inst.csv python tools/profile_sweep.py c3 151552 > /dev/null 2>&1
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:gsweep_kernel -s 1 -c 1 --csv --log-file gpurun_out/g_c5_inst.csv python tools/profile_gsweep.py 284160 10000 > /dev/null 2>&1
cat gpurun_out/g_c3.log gpurun_out/g_c5.log gpurun_out/g_tests.log; tail -n 3 gpurun_out/g_c3_inst.csv; tail -n 3 gpurun_out/g_c5_inst.csv
