#!/bin/bash
# reduced multi-GPU pass (one 8-GPU box): sharded verification at N=1 and N=8, batched registration at N=8
set -u
out=gpurun_out/scale2
mkdir -p $out
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@"; }
python bench.py --workload verify --steps 2 > $out/verify_1.json 2> $out/verify_1.err
tr 8 --workload verify --steps 2 > $out/verify_8.json 2> $out/verify_8.err
tr 8 --steps 20 > $out/align_8.json 2> $out/align_8.err
nvidia-smi -L > $out/gpus.txt
tail -n 2 $out/*.json
