#!/bin/bash
# round-2 call S: weight planes packed ahead on their own stream
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2s_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2s_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline"
$B --profile-out $O/r2s_cfg4_per_op.json > $O/r2s_bench.json 2> $O/r2s_bench.err
NLT_PACK_AHEAD=0 $B --no-parity > $O/r2s_bench_nopack.json 2> $O/r2s_bench_nopack.err
$B --no-graph --no-parity > $O/r2s_bench_eager.json 2> $O/r2s_bench_eager.err
timeout 300 python tools/opbench.py --graph > $O/r2s_graph_all.txt 2>&1
tail -2 $O/r2s_pytest.log; grep -E "^FAILED" $O/r2s_pytest.log | head; python -c "
import json
for f in ('r2s_bench','r2s_bench_nopack','r2s_bench_eager'):
    try:
        d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], (d.get('parity') or {}).get('ok'), d['roofline']['top5'][:3])
    except Exception as e: print(f, 'ERR', e)"
cat $O/r2s_graph_all.txt | tail -34
