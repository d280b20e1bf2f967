#!/bin/bash
# round-2 call W: consolidation -- full GPU suite, smoke, default bench (sub-records, parity gate, cpu baseline), per-layer
# graph-replay times, ncu --set full of the two level-0 kernels, ncu launch list of one eager step
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2w_pytest.log 2>&1
echo "gpu tests rc=$?" >> $O/r2w_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2w_smoke.log 2>&1
echo "smoke rc=$?" >> $O/r2w_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 --profile-out $O/r2w_cfg4_per_op.json > $O/r2w_bench.json 2> $O/r2w_bench.err
timeout 300 python tools/opbench.py --graph > $O/r2w_graph_all.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:pwx_wgrad_kernel -s 1 -c 1 -o $O/r2w_pwx_wgrad python tools/opbench.py --layers query.0.0 --cq-segs 3 60 1 --iters 2 --warmup 1 > $O/r2w_ncu1.log 2>&1
timeout 600 $NCU -k regex:pwx_fwd_kernel -s 1 -c 1 -o $O/r2w_pwx_fwd python tools/opbench.py --layers query.0.0 --cq-segs 3 60 1 --iters 2 --warmup 1 > $O/r2w_ncu2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r2w_ncu_launches.csv python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline --no-parity --no-graph > $O/r2w_ncu_bench.log 2>&1
ls -la $O/r2w_*.ncu-rep $O/r2w_ncu_launches.csv
tail -3 $O/r2w_pytest.log; grep -E "^FAILED" $O/r2w_pytest.log | head; tail -2 $O/r2w_smoke.log
python -c "
import json
d=json.loads(open('$O/r2w_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['step']['hbm_frac'], d['parity']['ok'], d['cpu_baseline']['value'])
for x in d.get('extra', []): print(x.get('workload'), x.get('batch_per_gpu'), x.get('ms_per_step'))
"
