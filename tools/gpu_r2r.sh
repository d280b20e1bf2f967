#!/bin/bash
# round-2 call R: pf kernel as up-conv input gradient, one-wave column tiling of the tcgen05 conv kernel, several weight-gradient streams
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2r_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2r_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline"
$B --profile-out $O/r2r_cfg4_per_op.json > $O/r2r_bench.json 2> $O/r2r_bench.err
NLT_TC_SPLIT_WAVES=1 $B --no-parity > $O/r2r_bench_splitwaves.json 2> $O/r2r_bench_splitwaves.err
NLT_SIDE_STREAMS=2 $B > $O/r2r_bench_side2.json 2> $O/r2r_bench_side2.err
NLT_SIDE_STREAMS=3 $B --no-parity > $O/r2r_bench_side3.json 2> $O/r2r_bench_side3.err
NLT_SIDE_STREAMS=4 $B --no-parity > $O/r2r_bench_side4.json 2> $O/r2r_bench_side4.err
timeout 300 python tools/opbench.py --graph > $O/r2r_graph_all.txt 2>&1
tail -2 $O/r2r_pytest.log; grep -E "^FAILED" $O/r2r_pytest.log | head; python -c "
import json
for f in ('r2r_bench','r2r_bench_splitwaves','r2r_bench_side2','r2r_bench_side3','r2r_bench_side4'):
    try:
        d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d.get('parity',{}).get('ok'), d['roofline']['top5'][:3])
    except Exception as e: print(f, 'ERR', e)"
cat $O/r2r_graph_all.txt | tail -34
