#!/bin/bash
# round-2 call Y: pwd2s with compile-time table offsets (32-channel sources), optionally ahead of the tensor path at level 3
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2y_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2y_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline"
$B --profile-out $O/r2y_cfg4_per_op.json > $O/r2y_bench.json 2> $O/r2y_bench.err
NLT_PWD2S_FIRST=1 $B --profile-out $O/r2y_cfg4_per_op_first.json > $O/r2y_bench_first.json 2> $O/r2y_bench_first.err
$B --no-parity > $O/r2y_bench_b.json 2> $O/r2y_bench_b.err
NLT_PWD2S_FIRST=1 $B --no-parity > $O/r2y_bench_first_b.json 2> $O/r2y_bench_first_b.err
tail -2 $O/r2y_pytest.log; grep -E "^FAILED" $O/r2y_pytest.log | head; python -c "
import json,glob
for f in sorted(glob.glob('$O/r2y_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), (d.get('parity') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)"
