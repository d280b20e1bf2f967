#!/bin/bash
# round-2 call H: (1) validate the pwd2s kernel + norm test, (2) ablation timing of the tcgen05 conv kernel on mid/deep layers
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_norm.py tests/test_gpu_parity.py tests/test_gpu_model.py -m gpu -q > $O/r2h_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2h_pytest.log
LAYERS="query.3.0 query.4.0 query.4.1 query.6.0 query.8.0"
for a in 0 1 16 17 2 4 8 31; do
  NLT_TC_ABLATE=$a timeout 200 python tools/opbench.py --layers $LAYERS --iters 5 > $O/r2h_ablate_$a.txt 2>&1
done
timeout 300 python tools/opbench.py --layers query.1.0 obs.1.0 query.2.0 obs.2.0 > $O/r2h_opbench_pwd.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2h_cfg4_per_op.json > $O/r2h_bench.json 2> $O/r2h_bench.err
tail -3 $O/r2h_pytest.log; for a in 0 1 16 17 2 4 8 31; do echo "ablate $a"; tail -5 $O/r2h_ablate_$a.txt; done; cat $O/r2h_opbench_pwd.txt | tail -4; head -c 400 $O/r2h_bench.json
