#!/bin/bash
# round-2 call P: staged-patch forward of the stride-2 2x2 convs (pf_fwd)
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2p_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2p_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2p_cfg4_per_op.json > $O/r2p_bench.json 2> $O/r2p_bench.err
NLT_PF=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity > $O/r2p_bench_nopf.json 2> $O/r2p_bench_nopf.err
timeout 300 python tools/opbench.py --graph > $O/r2p_graph_all.txt 2>&1
tail -2 $O/r2p_pytest.log; grep -E "^FAILED" $O/r2p_pytest.log | head; python -c "
import json
for f in ('r2p_bench','r2p_bench_nopf'):
    try:
        d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['top5'][:4])
    except Exception as e: print(f, 'ERR', e)"
cat $O/r2p_graph_all.txt | tail -34
