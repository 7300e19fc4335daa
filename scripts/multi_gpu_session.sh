#!/bin/bash
# One multi-GPU session: NCCL correctness check, configs 3 / 4 / 5, the bench line with its sharded leg.
# usage: scripts/multi_gpu_session.sh N [small]
N=${1:-2}
MODE=${2:-full}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_mg_N${N}.log 2>&1
echo "== multi_gpu_check" >> gpurun_out/r2_mg_N${N}.log
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py >> gpurun_out/r2_mg_N${N}.log 2>&1; echo "EXIT $?" >> gpurun_out/r2_mg_N${N}.log
if [ "$MODE" = "small" ]; then
  A3="--bursts 24 --wave 12"; A4="--grid 8"; A5="--nfreq5 256 --ntime5 16384"
else
  A3="--bursts 4096 --wave 48"; A4="--grid 1024"; A5=""
fi
for cfg in 3 5 4; do
  echo "== config $cfg" >> gpurun_out/r2_mg_N${N}.log
  eval "ARGS=\$A$cfg"
  timeout 900 $TR --master-port 2953$cfg scripts/run_configs_multi.py $cfg $ARGS >> gpurun_out/r2_mg_N${N}.log 2>&1; echo "EXIT $?" >> gpurun_out/r2_mg_N${N}.log
done
echo "== bench" >> gpurun_out/r2_mg_N${N}.log
timeout 900 $TR --master-port 29540 bench.py --gpus $N --steps 24 --warmup 8 > gpurun_out/r2_bench_N${N}.json 2> gpurun_out/r2_bench_N${N}.err; echo "EXIT $?" >> gpurun_out/r2_mg_N${N}.log
grep -v "^\s*$" gpurun_out/r2_mg_N${N}.log | grep -E "EXIT|==|MULTI_GPU_OK|Error|error|\"config\"" | cut -c1-400
