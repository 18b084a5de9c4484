#!/bin/bash
# Multi-GPU measurement suite (one box, 8 GPUs): weak scaling of the bench, strong scaling of config 5 (independent-agent form)
# and of config 4 (ADMM consensus, NCCL all-gather per round).  Output: gpurun_out/multigpu_<tag>.log (one JSON line per run).
tag=${1:-r01h}
out=gpurun_out/multigpu_${tag}.log
mkdir -p gpurun_out; : > $out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 2000)) "${@:2}"; }
echo "# bench weak scaling" >> $out
for n in 2 4 8; do tr $n bench.py --gpus $n --steps 10 --warmup 3 2>/dev/null | tail -1 >> $out; done
echo "# config5 strong scaling (8192 agents x K=200, M=32)" >> $out
python tools/run_config5.py 8192 3 2>/dev/null | tail -1 >> $out
for n in 2 4 8; do tr $n tools/run_config5.py 8192 3 2>/dev/null | tail -1 >> $out; done
echo "# config4 strong scaling (256 SI agents, all pairs, 10 ADMM rounds)" >> $out
python tools/run_admm_multi.py 4 2>/dev/null | tail -1 >> $out
for n in 2 4 8; do tr $n tools/run_admm_multi.py 4 2>/dev/null | tail -1 >> $out; done
cut -c1-260 $out
