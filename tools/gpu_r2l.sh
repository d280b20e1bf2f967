#!/bin/bash
# round-2 call L: generalised pwd2s (32-channel sources, multi-row tiles) -- tests, bench
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2l_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2l_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2l_cfg4_per_op.json > $O/r2l_bench.json 2> $O/r2l_bench.err
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2l_cfg2_per_op.json > $O/r2l_bench_cfg2.json 2> $O/r2l_bench_cfg2.err
tail -2 $O/r2l_pytest.log; python -c "
import json
for f in ('r2l_bench','r2l_bench_cfg2'):
    d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['top5'][:4])"
