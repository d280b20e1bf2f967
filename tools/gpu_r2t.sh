#!/bin/bash
# round-2 call T: graph vs eager scheduling matrix; threads-per-pixel variants of the FFMA2 forward kernels
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q > $O/r2t_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2t_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity"
run() { tag=$1; shift; env "$@" $B $EXTRA > $O/r2t_bench_$tag.json 2> $O/r2t_bench_$tag.err; }
EXTRA=""
run g_ns22 X=1
run g_ns11 NLT_PWX_NS=1 NLT_PF_NS=1
run g_ns12 NLT_PWX_NS=1 NLT_PF_NS=2
run g_ns21 NLT_PWX_NS=2 NLT_PF_NS=1
EXTRA="--no-graph"
run e_default X=1
run e_nopack NLT_PACK_AHEAD=0
run e_side1 NLT_SIDE_STREAMS=1
run e_side1_nopack NLT_SIDE_STREAMS=1 NLT_PACK_AHEAD=0
run e_side3 NLT_SIDE_STREAMS=3
run e_noside NLT_NO_SIDE_STREAM=1
EXTRA=""
run g_side1_nopack NLT_SIDE_STREAMS=1 NLT_PACK_AHEAD=0
$B --profile-out $O/r2t_cfg4_per_op.json > $O/r2t_bench_prof.json 2> $O/r2t_bench_prof.err
tail -2 $O/r2t_pytest.log; grep -E "^FAILED" $O/r2t_pytest.log | head; python -c "
import json,glob
for f in sorted(glob.glob('$O/r2t_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('r2t_bench_')[1], round(d['ms_per_step'],3), d.get('whole_step_in_cuda_graph'))
    except Exception as e: print(f, 'ERR', e)"
