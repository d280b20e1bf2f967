#!/bin/bash
# round-2 call Z: ablation of the tcgen05 conv kernel on the many-tile mid-level layers (levels 2-3), graph replay
O=gpurun_out
MID="obs.2.1 query.2.1 query.2.0 obs.3.0 query.3.0 query.3.1 query.4.0 query.4.1"
timeout 200 python tools/opbench.py --graph --layers $MID > $O/r2z_mid_full.txt 2>&1
for a in 1 2 4 8 16 17 31; do
  NLT_TC_ABLATE=$a timeout 200 python tools/opbench.py --graph --layers $MID > $O/r2z_mid_ablate_$a.txt 2>&1
done
NLT_DISABLE_TC=1 timeout 300 python tools/opbench.py --graph --layers $MID > $O/r2z_mid_notc.txt 2>&1
echo full; tail -8 $O/r2z_mid_full.txt; for a in 1 2 4 8 16 17 31; do echo "ablate $a"; tail -8 $O/r2z_mid_ablate_$a.txt; done; echo notc; tail -8 $O/r2z_mid_notc.txt
