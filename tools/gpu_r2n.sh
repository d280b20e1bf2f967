#!/bin/bash
# round-2 call N: consolidation -- whole GPU suite, smoke, default bench line with every sub-record, GPU-only layer table
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2n_pytest.log 2>&1
echo "gpu tests rc=$?" >> $O/r2n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2n_smoke.log 2>&1
timeout 1200 python bench.py --steps 20 --warmup 5 --profile-out $O/r2n_cfg4_per_op.json > $O/r2n_bench.json 2> $O/r2n_bench.err
echo "bench rc=$?" >> $O/r2n_bench.err
ALL="obs.0.0 query.0.0 obs.1.0 query.1.0 obs.1.1 query.1.1 obs.2.0 query.2.0 obs.2.1 query.2.1 obs.3.0 query.3.0 query.3.1 query.4.0 query.4.1 query.5.0 query.5.1 query.6.0 query.6.1 query.7.0 query.7.1 query.8.0 query.8.1 query.9.0 query.9.1 query.10.0 query.10.1 query.11.0 query.11.1 query.12.0 query.12.1 query.13.0"
timeout 600 python tools/opbench.py --graph --cq-segs 3 60 1 --layers $ALL > $O/r2n_graph_all.txt 2>&1
tail -3 $O/r2n_pytest.log; tail -1 $O/r2n_smoke.log; tail -2 $O/r2n_bench.err; cat $O/r2n_graph_all.txt
python -c "
import json
d=json.loads(open('$O/r2n_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['uint8_inputs']['ms_per_step'], d['roofline']['frac'], d['roofline']['step']['hbm_frac'])
for e in d['extra']: print(e.get('workload'), e.get('batch_per_gpu'), e.get('ms_per_step'), e.get('error'))"
