#!/bin/bash
# round-2 call AF: source-level ncu of the tcgen05 conv kernel on a many-tile mid-level layer (real and fully ablated)
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:tc_gconv_kernel -s 2 -c 1 -o $O/r2af_tc_q21 python tools/opbench.py --layers query.2.1 --iters 2 --warmup 1 > $O/r2af_ncu1.log 2>&1
NLT_TC_ABLATE=31 timeout 300 $NCU -k regex:tc_gconv_kernel -s 2 -c 1 -o $O/r2af_tc_q21_ablate31 python tools/opbench.py --layers query.2.1 --iters 2 --warmup 1 > $O/r2af_ncu2.log 2>&1
ls -la $O/r2af_*.ncu-rep; tail -2 $O/r2af_ncu1.log
