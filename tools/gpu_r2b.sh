#!/bin/bash
# round-2 call B: full GPU suite on the new host code + the pwx kernels, level-0 A/B at the cfg4 shape
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2b_pytest.log 2>&1
echo "gpu tests rc=$?" >> $O/r2b_pytest.log
timeout 300 python tools/opbench.py --layers query.0.0 --cq-segs 3 60 1 > $O/r2b_opbench_pwx1.txt 2>&1
timeout 300 python tools/opbench.py --layers query.0.0 --cq-segs 3 60 1 --opt pwx=0 > $O/r2b_opbench_pwx0.txt 2>&1
timeout 600 python bench.py --c-extra 59 --steps 10 --warmup 3 --no-cpu-baseline --profile-out $O/r2b_cfg4_b8_per_op.json > $O/r2b_cfg4_b8.json 2> $O/r2b_cfg4_b8.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out $O/r2b_cfg2_per_op.json > $O/r2b_cfg2.json 2> $O/r2b_cfg2.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2b_smoke.log 2>&1
tail -4 $O/r2b_pytest.log; cat $O/r2b_opbench_pwx1.txt $O/r2b_opbench_pwx0.txt | tail -6; tail -2 $O/r2b_smoke.log
head -c 300 $O/r2b_cfg4_b8.json; tail -c 600 $O/r2b_cfg4_b8.err
