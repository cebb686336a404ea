#!/bin/bash
# batched registration at N=8 (weak scaling), device-resident and end to end
set -u
mkdir -p gpurun_out/scale3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 > gpurun_out/scale3/align_8.json 2> gpurun_out/scale3/align_8.err
tail -n 1 gpurun_out/scale3/align_8.json
nvidia-smi topo -m > gpurun_out/scale3/topo.txt 2>&1
