#!/bin/bash
# round-2 call Q: outer-product weight gradients (nlt_wop.cu) + ncu of the FFMA2 stream kernels
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2q_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2q_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2q_cfg4_per_op.json > $O/r2q_bench.json 2> $O/r2q_bench.err
NLT_WOP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity --profile-out $O/r2q_cfg4_per_op_nowop.json > $O/r2q_bench_nowop.json 2> $O/r2q_bench_nowop.err
timeout 300 python tools/opbench.py --graph > $O/r2q_graph_all.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:pwx_fwd_kernel -s 1 -c 1 -o $O/r2q_pwx_fwd python tools/opbench.py --layers query.0.0 --cq-segs 3 60 1 --iters 2 --warmup 1 > $O/r2q_ncu1.log 2>&1
timeout 600 $NCU -k regex:pf_fwd_kernel -s 1 -c 1 -o $O/r2q_pf_fwd python tools/opbench.py --layers query.1.0 --iters 2 --warmup 1 > $O/r2q_ncu2.log 2>&1
timeout 600 $NCU -k regex:tiny_stencil_kernel -s 2 -c 2 -o $O/r2q_tiny python tools/opbench.py --layers query.1.1 --iters 2 --warmup 1 > $O/r2q_ncu3.log 2>&1
timeout 600 $NCU -k regex:wop_wgrad_kernel -s 1 -c 1 -o $O/r2q_wop_q11 python tools/opbench.py --layers query.1.1 --iters 2 --warmup 1 > $O/r2q_ncu4.log 2>&1
timeout 600 $NCU -k regex:wop_wgrad_kernel -s 1 -c 1 -o $O/r2q_wop_q120 python tools/opbench.py --layers query.12.0 --iters 2 --warmup 1 > $O/r2q_ncu5.log 2>&1
timeout 600 $NCU -k regex:pwx_d2s_fwd_kernel -s 1 -c 1 -o $O/r2q_d2s_q110 python tools/opbench.py --layers query.11.0 --iters 2 --warmup 1 > $O/r2q_ncu6.log 2>&1
ls -la $O/*.ncu-rep
tail -2 $O/r2q_pytest.log; grep -E "^FAILED" $O/r2q_pytest.log | head; python -c "
import json
for f in ('r2q_bench','r2q_bench_nowop'):
    try:
        d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['top5'][:4])
    except Exception as e: print(f, 'ERR', e)"
cat $O/r2q_graph_all.txt | tail -34
