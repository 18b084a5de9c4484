out=gpurun_out/config5_admm_r01i.log; : > $out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 2000)) "${@:2}"; }
python tools/run_config5_admm.py 8192 3 16 2>/dev/null | tail -1 >> $out
for n in 2 4 8; do tr $n tools/run_config5_admm.py 8192 3 16 2>/dev/null | tail -1 >> $out; done
cut -c1-420 $out
