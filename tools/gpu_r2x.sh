#!/bin/bash
# round-2 call X (2 GPUs): data-parallel step with the all-reduce inside the CUDA graph -- correctness, then the bench
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py > $O/r2x_dist_check.json 2> $O/r2x_dist_check.err
echo "dist_check rc=$?" >> $O/r2x_dist_check.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2x_bench_n2.json 2> $O/r2x_bench_n2.err
echo "bench n2 rc=$?" >> $O/r2x_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2x_ref_n2.json 2> $O/r2x_ref_n2.err
cat $O/r2x_dist_check.json; tail -3 $O/r2x_dist_check.err; head -c 600 $O/r2x_bench_n2.json; tail -3 $O/r2x_bench_n2.err
