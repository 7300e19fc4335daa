#!/bin/bash
# NCCL correctness check + bench line (with its sharded leg) on N GPUs: scripts/multi_gpu_check_and_bench.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
L=gpurun_out/r2_final_mg_N${N}.log
nvidia-smi -L > $L 2>&1
echo "== multi_gpu_check" >> $L
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py >> $L 2>&1; echo "EXIT $?" >> $L
echo "== bench" >> $L
timeout 900 $TR --master-port 29540 bench.py --gpus $N --steps 24 --warmup 8 > gpurun_out/r2_final_bench_N${N}.json 2> gpurun_out/r2_final_bench_N${N}.err; echo "EXIT $?" >> $L
grep -E "EXIT|==|MULTI_GPU_OK|Error" $L | cut -c1-300
