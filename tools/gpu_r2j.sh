#!/bin/bash
# round-2 call J: pwx kernels at 4 CTAs/SM, pwd2s v2 -- tests, GPU-only layer times, bench
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q > $O/r2j_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2j_pytest.log
timeout 300 python tools/opbench.py --graph --cq-segs 3 60 1 --layers query.0.0 obs.1.0 query.1.0 obs.2.0 query.2.0 > $O/r2j_graph.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2j_cfg4_per_op.json > $O/r2j_bench.json 2> $O/r2j_bench.err
tail -2 $O/r2j_pytest.log; cat $O/r2j_graph.txt; python -c "
import json
d=json.loads(open('$O/r2j_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['top5'])"
