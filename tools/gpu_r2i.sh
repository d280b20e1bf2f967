#!/bin/bash
# round-2 call I: GPU-only (graph replay) layer times; mbarrier polling / ablation of the tcgen05 skeleton; pwd2s v2
O=gpurun_out
ALL="obs.0.0 query.0.0 obs.1.0 query.1.0 obs.1.1 query.1.1 obs.2.0 query.2.0 obs.2.1 query.2.1 obs.3.0 query.3.0 query.3.1 query.4.0 query.4.1 query.5.0 query.5.1 query.6.0 query.6.1 query.7.0 query.7.1 query.8.0 query.8.1 query.9.0 query.9.1 query.10.0 query.10.1 query.11.0 query.11.1 query.12.0 query.12.1 query.13.0"
timeout 600 python tools/opbench.py --graph --cq-segs 3 60 1 --layers $ALL > $O/r2i_graph_all.txt 2>&1
DEEP="query.3.0 query.4.0 query.4.1 query.5.0 query.6.0 query.7.0 query.8.0 query.9.0"
for a in 31 32 63 4 2 17; do
  NLT_TC_ABLATE=$a timeout 200 python tools/opbench.py --graph --layers $DEEP > $O/r2i_graph_ablate_$a.txt 2>&1
done
NLT_DISABLE_TC=1 timeout 300 python tools/opbench.py --graph --layers $DEEP > $O/r2i_graph_notc.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -m gpu -q -k "rides_on or forward_backward or full_size" > $O/r2i_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2i_pytest.log
cat $O/r2i_graph_all.txt; for a in 31 32 63 4 2 17; do echo "ablate $a"; tail -8 $O/r2i_graph_ablate_$a.txt; done; echo notc; tail -8 $O/r2i_graph_notc.txt; tail -2 $O/r2i_pytest.log
