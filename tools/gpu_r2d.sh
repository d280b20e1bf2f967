#!/bin/bash
# round-2 call D: TS-form tcgen05 kernel -- correctness (both forms), per-layer A/B, whole step with NLT_TCS=1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -k "tensor_core or k_observations" > $O/r2d_pytest_ts.log 2>&1
echo "ts tests rc=$?" >> $O/r2d_pytest_ts.log
LAYERS="query.1.0 obs.1.0 query.1.1 obs.2.0 query.2.0 query.2.1 query.3.0 query.3.1 query.4.0 query.4.1 query.5.0 query.6.0 query.7.0 query.8.0 query.9.0 query.10.0 query.10.1 query.11.0"
timeout 300 python tools/opbench.py --layers $LAYERS > $O/r2d_opbench_ss.txt 2>&1
NLT_TCS=1 timeout 300 python tools/opbench.py --layers $LAYERS > $O/r2d_opbench_ts64.txt 2>&1
NLT_TCS=1 NLT_TCS_KMIN=16 timeout 300 python tools/opbench.py --layers $LAYERS > $O/r2d_opbench_ts16.txt 2>&1
NLT_TCS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2d_cfg4_ts64_per_op.json > $O/r2d_bench_ts64.json 2> $O/r2d_bench_ts64.err
NLT_TCS=1 NLT_TCS_KMIN=16 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2d_cfg4_ts16_per_op.json > $O/r2d_bench_ts16.json 2> $O/r2d_bench_ts16.err
NLT_TCS=1 timeout 1200 python -m pytest tests -m gpu -q > $O/r2d_pytest_all_ts.log 2>&1
echo "all tests (TS) rc=$?" >> $O/r2d_pytest_all_ts.log
tail -3 $O/r2d_pytest_ts.log; paste $O/r2d_opbench_ss.txt $O/r2d_opbench_ts64.txt $O/r2d_opbench_ts16.txt | cut -c1-200; tail -2 $O/r2d_pytest_all_ts.log
python -c "
import json
for f in ('ts64','ts16'):
    try:
        d=json.loads(open('$O/r2d_bench_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['top5'])
    except Exception as e: print(f,'ERR',e)
"
