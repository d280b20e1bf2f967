#!/bin/bash
# round-2 call F: suite on the pipelined tc epilogue / constant-bank pwx forward / uint8 inputs / norm layers; layer A/B; bench
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2f_pytest.log 2>&1
echo "gpu tests rc=$?" >> $O/r2f_pytest.log
LAYERS="query.1.0 obs.1.0 query.1.1 obs.2.0 query.2.0 query.2.1 query.3.0 query.3.1 query.4.0 query.4.1 query.5.0 query.6.0 query.7.0 query.8.0 query.9.0 query.10.0 query.10.1 query.11.0"
timeout 300 python tools/opbench.py --layers $LAYERS > $O/r2f_opbench_ss.txt 2>&1
timeout 300 python tools/opbench.py --layers query.0.0 --cq-segs 3 60 1 > $O/r2f_opbench_q00.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out $O/r2f_cfg4_per_op.json > $O/r2f_bench.json 2> $O/r2f_bench.err
echo "bench rc=$?" >> $O/r2f_bench.err
tail -4 $O/r2f_pytest.log; paste $O/r2d_opbench_ss.txt $O/r2f_opbench_ss.txt 2>/dev/null | cut -c1-170; cat $O/r2f_opbench_q00.txt | tail -2; head -c 900 $O/r2f_bench.json; tail -2 $O/r2f_bench.err
