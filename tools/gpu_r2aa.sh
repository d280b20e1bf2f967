#!/bin/bash
# round-2 call AA: coalesced weight gradient of the final 1x1 conv (wopn), pwd2s ahead of the tensor path at level 3
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2aa_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2aa_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline"
$B --profile-out $O/r2aa_cfg4_per_op.json > $O/r2aa_bench.json 2> $O/r2aa_bench.err
NLT_WOP=0 $B --no-parity > $O/r2aa_bench_nowop.json 2> $O/r2aa_bench_nowop.err
tail -2 $O/r2aa_pytest.log; grep -E "^FAILED" $O/r2aa_pytest.log | head; python -c "
import json,glob
for f in sorted(glob.glob('$O/r2aa_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), (d.get('parity') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)
a=json.load(open('$O/r2aa_cfg4_per_op.json'))
for r in a['rows']:
    if 'query.13.0' in r['op'] or 'query.3.0' in r['op'] or 'obs.3.0' in r['op']: print(r['op'], round(r['ms_per_step'],3))
"
