#!/bin/bash
# round-2 call M: pws weight gradient, pwd2s ahead of the tensor path, index shifts -- tests, bench, no-wgrad timing
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity.py -m gpu -q > $O/r2m_pytest.log 2>&1
echo "tests rc=$?" >> $O/r2m_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --profile-out $O/r2m_cfg4_per_op.json > $O/r2m_bench.json 2> $O/r2m_bench.err
NLT_SKIP_WGRAD=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity > $O/r2m_bench_nowgrad.json 2> $O/r2m_bench_nowgrad.err
NLT_PWX=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-parity > $O/r2m_bench_nopwx.json 2> $O/r2m_bench_nopwx.err
tail -2 $O/r2m_pytest.log; grep -E "^FAILED" $O/r2m_pytest.log | head; python -c "
import json
for f in ('r2m_bench','r2m_bench_nowgrad','r2m_bench_nopwx'):
    try:
        d=json.loads(open('$O/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['top5'][:4])
    except Exception as e: print(f, 'ERR', e)"
