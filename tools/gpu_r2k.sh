#!/bin/bash
# round-2 call K: one-pass depth-to-space tcgen05 weight gradient; what the side stream buys inside the graph
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2k_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2k_pytest.log
timeout 300 python tools/opbench.py --graph --layers query.7.0 query.8.0 query.9.0 query.10.0 > $O/r2k_graph_up.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2k_cfg4_per_op.json > $O/r2k_bench.json 2> $O/r2k_bench.err
NLT_NO_SIDE_STREAM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity > $O/r2k_bench_noside.json 2> $O/r2k_bench_noside.err
tail -2 $O/r2k_pytest.log; cat $O/r2k_graph_up.txt; python -c "
import json
for f in ('r2k_bench','r2k_bench_noside'):
    d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['top5'][:3])"
