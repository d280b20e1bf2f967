#!/bin/bash
# round-2 call AH: raw activation tile as the hi operand of the tcgen05 conv kernel (option tc_rawhi)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "raw_hi or tensor_core" > $O/r2ah_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2ah_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline"
NLT_TC_RAWHI=1 $B > $O/r2ah_bench_rawhi.json 2> $O/r2ah_bench_rawhi.err
$B --no-parity > $O/r2ah_bench.json 2> $O/r2ah_bench.err
NLT_TC_RAWHI=1 timeout 300 python tools/opbench.py --graph > $O/r2ah_graph_all_rawhi.txt 2>&1
tail -2 $O/r2ah_pytest.log; grep -E "^FAILED" $O/r2ah_pytest.log | head; python -c "
import json,glob
for f in sorted(glob.glob('$O/r2ah_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), (d.get('parity') or {}))
    except Exception as e: print(f, 'ERR', e)"
grep -E "^(query|obs)\.[2-9]\." $O/r2ah_graph_all_rawhi.txt
