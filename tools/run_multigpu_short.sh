#!/bin/bash
# Reduced 8-GPU suite (after a kernel change): bench weak scaling at 8, config 5 (independent form) at 1 and 8, config 5
# (decentralised) at 8.  Output: gpurun_out/multigpu_short_<tag>.log
tag=${1:-r01j}
out=gpurun_out/multigpu_short_${tag}.log
mkdir -p gpurun_out; : > $out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 2000)) "${@:2}"; }
tr 8 bench.py --gpus 8 --steps 20 --warmup 3 2>/dev/null | tail -1 >> $out
python tools/run_config5.py 8192 3 2>/dev/null | tail -1 >> $out
tr 8 tools/run_config5.py 8192 3 2>/dev/null | tail -1 >> $out
tr 8 tools/run_config5_admm.py 8192 3 16 2>/dev/null | tail -1 >> $out
cut -c1-330 $out
