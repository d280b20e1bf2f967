#!/bin/bash
# round-2 call C: whole GPU suite (new parity tests, TS probe, deterministic scatter), the new default bench line, launch list
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2c_pytest.log 2>&1
echo "gpu tests rc=$?" >> $O/r2c_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out $O/r2c_cfg4_per_op.json > $O/r2c_bench.json 2> $O/r2c_bench.err
echo "bench rc=$?" >> $O/r2c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2c_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-extra --no-cpu-baseline --no-parity > $O/r2c_ncu_bench.log 2>&1
tail -5 $O/r2c_pytest.log; head -c 1500 $O/r2c_bench.json; tail -3 $O/r2c_bench.err; ls gpurun_out | grep -c parity
