#!/bin/bash
# round-2 call U: threads-per-pixel variants of the FFMA2 forward kernels (compile-time weight column offsets)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q > $O/r2u_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2u_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity"
run() { tag=$1; shift; env "$@" $B $EXTRA > $O/r2u_bench_$tag.json 2> $O/r2u_bench_$tag.err; }
run ns22 X=1
run ns11 NLT_PWX_NS=1 NLT_PF_NS=1
run ns12 NLT_PWX_NS=1 NLT_PF_NS=2
run ns21 NLT_PWX_NS=2 NLT_PF_NS=1
run ns11_b NLT_PWX_NS=1 NLT_PF_NS=1
run ns22_b X=1
$B --profile-out $O/r2u_cfg4_per_op.json > $O/r2u_bench_prof.json 2> $O/r2u_bench_prof.err
tail -2 $O/r2u_pytest.log; grep -E "^FAILED" $O/r2u_pytest.log | head; python -c "
import json,glob
for f in sorted(glob.glob('$O/r2u_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('r2u_bench_')[1], round(d['ms_per_step'],3), d.get('whole_step_in_cuda_graph'))
    except Exception as e: print(f, 'ERR', e)"
