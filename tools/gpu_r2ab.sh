#!/bin/bash
# round-2 call AB: neighbour-merged fixed-point scatter of the tail; staged-patch kernel on the stride-1 16 -> 16 stencils
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2ab_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2ab_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline"
$B --profile-out $O/r2ab_cfg4_per_op.json > $O/r2ab_bench.json 2> $O/r2ab_bench.err
NLT_PF_S1=1 $B --profile-out $O/r2ab_cfg4_per_op_s1.json > $O/r2ab_bench_s1.json 2> $O/r2ab_bench_s1.err
tail -2 $O/r2ab_pytest.log; grep -E "^FAILED" $O/r2ab_pytest.log | head; python -c "
import json,glob
for f in sorted(glob.glob('$O/r2ab_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), (d.get('parity') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)
a=json.load(open('$O/r2ab_cfg4_per_op.json')); b=json.load(open('$O/r2ab_cfg4_per_op_s1.json'))
bm={r['op']:r['ms_per_step'] for r in b['rows']}
for r in a['rows']:
    if 'tail' in r['op'] or '.1.1' in r['op'] or '10.1' in r['op']: print(r['op'], round(r['ms_per_step'],3), 's1', round(bm[r['op']],3))
"
