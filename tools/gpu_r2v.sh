#!/bin/bash
# round-2 call V: pointwise FFMA2 weight gradient for the up-convs into 4 / 8 channels (depth-to-space gather of the gradient)
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2v_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2v_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline"
$B --profile-out $O/r2v_cfg4_per_op.json > $O/r2v_bench.json 2> $O/r2v_bench.err
$B --no-parity > $O/r2v_bench_b.json 2> $O/r2v_bench_b.err
timeout 300 python tools/opbench.py --graph > $O/r2v_graph_all.txt 2>&1
tail -2 $O/r2v_pytest.log; grep -E "^FAILED" $O/r2v_pytest.log | head; python -c "
import json,glob
for f in sorted(glob.glob('$O/r2v_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), (d.get('parity') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)"
