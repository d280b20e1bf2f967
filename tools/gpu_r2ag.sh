#!/bin/bash
# round-2 call AG: leaner epilogue of the tcgen05 conv kernel (separable output offsets, float4 bias, hoisted activation branch)
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2ag_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2ag_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline"
$B --profile-out $O/r2ag_cfg4_per_op.json > $O/r2ag_bench.json 2> $O/r2ag_bench.err
timeout 300 python tools/opbench.py --graph > $O/r2ag_graph_all.txt 2>&1
NLT_TC_ABLATE=31 timeout 200 python tools/opbench.py --graph --layers obs.2.1 query.2.1 query.2.0 obs.3.0 query.3.0 query.3.1 query.4.0 query.4.1 > $O/r2ag_mid_ablate_31.txt 2>&1
tail -2 $O/r2ag_pytest.log; grep -E "^FAILED" $O/r2ag_pytest.log | head; python -c "
import json,glob
for f in sorted(glob.glob('$O/r2ag_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), (d.get('parity') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)"
tail -33 $O/r2ag_graph_all.txt; echo ablate31; tail -8 $O/r2ag_mid_ablate_31.txt
