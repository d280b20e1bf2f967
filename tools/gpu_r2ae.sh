#!/bin/bash
# round-2 call AE: final consolidation -- full GPU suite, smoke, default bench (sub-records, parity gate, cpu baseline),
# per-layer graph-replay times
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2ae_pytest.log 2>&1
echo "gpu tests rc=$?" >> $O/r2ae_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2ae_smoke.log 2>&1
echo "smoke rc=$?" >> $O/r2ae_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 --profile-out $O/r2ae_cfg4_per_op.json > $O/r2ae_bench.json 2> $O/r2ae_bench.err
timeout 300 python tools/opbench.py --graph > $O/r2ae_graph_all.txt 2>&1
tail -3 $O/r2ae_pytest.log; grep -E "^FAILED" $O/r2ae_pytest.log | head; tail -2 $O/r2ae_smoke.log
python -c "
import json
d=json.loads(open('$O/r2ae_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['step']['hbm_frac'], d['parity']['ok'], d['cpu_baseline']['value'])
for x in d.get('extra', []): print(x.get('workload'), x.get('batch_per_gpu'), x.get('ms_per_step'))
"
