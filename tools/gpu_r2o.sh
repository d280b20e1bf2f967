#!/bin/bash
# round-2 call O: row-stream stencil kernel (nlt_tiny.cu) + depth-to-space up-conv forward (pwx_d2s_fwd)
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2o_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2o_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2o_cfg4_per_op.json > $O/r2o_bench.json 2> $O/r2o_bench.err
NLT_TINY=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity > $O/r2o_bench_notiny.json 2> $O/r2o_bench_notiny.err
timeout 300 python tools/opbench.py --graph > $O/r2o_graph_all.txt 2>&1
tail -2 $O/r2o_pytest.log; grep -E "^FAILED" $O/r2o_pytest.log | head; python -c "
import json
for f in ('r2o_bench','r2o_bench_notiny'):
    try:
        d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['top5'][:4])
    except Exception as e: print(f, 'ERR', e)"
cat $O/r2o_graph_all.txt | tail -34
