#!/bin/bash
# round-2 call A: validate the never-run kernels, first cfg4 (64-channel) measurements
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
NLT_TEST_EXPERIMENTAL=1 NLT_EXPERIMENTAL_BARRON=1 timeout 900 python -m pytest tests -m gpu -q -k "experimental or barron or wide" > $O/r2a_pytest_experimental.log 2>&1
echo "exp tests rc=$?" >> $O/r2a_pytest_experimental.log
timeout 600 python bench.py --c-extra 59 --steps 10 --warmup 3 --no-cpu-baseline --profile-out $O/r2a_cfg4_b8_per_op.json > $O/r2a_cfg4_b8.json 2> $O/r2a_cfg4_b8.err
timeout 600 python bench.py --c-extra 59 --batch 1 --steps 10 --warmup 3 --no-cpu-baseline --profile-out $O/r2a_cfg4_b1_per_op.json > $O/r2a_cfg4_b1.json 2> $O/r2a_cfg4_b1.err
timeout 600 python bench.py --c-extra 59 --batch 32 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2a_cfg4_b32.json 2> $O/r2a_cfg4_b32.err
for v in 1 2; do NLT_DCONV_WIDE32=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out $O/r2a_wide32_$v.json > $O/r2a_wide32_$v.bench.json 2> $O/r2a_wide32_$v.err; done
NLT_DCONV_WIDE_FIRST=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out $O/r2a_wide_first.json > $O/r2a_wide_first.bench.json 2>$O/r2a_wide_first.err
NLT_DCONV_WIDE_FIRST=1 NLT_DCONV_WIDE8=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out $O/r2a_wide8.json > $O/r2a_wide8.bench.json 2>$O/r2a_wide8.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out $O/r2a_cfg2_per_op.json > $O/r2a_cfg2.json 2> $O/r2a_cfg2.err
tail -3 $O/r2a_pytest_experimental.log
cat $O/r2a_cfg4_b8.json | head -c 600
