#!/bin/bash
# round-2 call E: source-level ncu of the tcgen05 conv kernels (SS and TS form) on a deep and a mid layer
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc.*_gconv_kernel -s 6 -c 3 -o $O/r2e_tc_ss_q40 python tools/opbench.py --layers query.4.0 --iters 2 --warmup 1 > $O/r2e_ncu1.log 2>&1
NLT_TCS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tcs_gconv_kernel -s 6 -c 3 -o $O/r2e_tc_ts_q40 python tools/opbench.py --layers query.4.0 --iters 2 --warmup 1 > $O/r2e_ncu2.log 2>&1
NLT_TCS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tcs_gconv_kernel -s 6 -c 3 -o $O/r2e_tc_ts_q11 python tools/opbench.py --layers query.1.1 --iters 2 --warmup 1 > $O/r2e_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pwx_fwd_kernel -s 1 -c 1 -o $O/r2e_pwx_fwd python tools/opbench.py --layers query.0.0 --cq-segs 3 60 1 --iters 2 --warmup 1 > $O/r2e_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_conv_kernel -s 2 -c 2 -o $O/r2e_pw_q10_dgrad python tools/opbench.py --layers query.1.0 --iters 2 --warmup 1 > $O/r2e_ncu5.log 2>&1
ls -la $O/*.ncu-rep; tail -2 $O/r2e_ncu1.log
