python tools/try_c3.py parity 1000000 2>&1 | tail -8
python tools/try_gsweep.py parity time 1000000 2>&1 | tail -4
