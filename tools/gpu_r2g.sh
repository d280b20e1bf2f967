#!/bin/bash
# round-2 call G (2 GPUs): data-parallel step with the all-reduce inside the CUDA graph -- correctness, then the bench
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > $O/r2g_dist_check.json 2> $O/r2g_dist_check.err
echo "dist_check rc=$?" >> $O/r2g_dist_check.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2g_bench_n2.json 2> $O/r2g_bench_n2.err
echo "bench n2 rc=$?" >> $O/r2g_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2g_ref_n2.json 2> $O/r2g_ref_n2.err
cat $O/r2g_dist_check.json; tail -3 $O/r2g_dist_check.err; head -c 600 $O/r2g_bench_n2.json; tail -3 $O/r2g_bench_n2.err
