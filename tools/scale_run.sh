#!/bin/bash
# multi-GPU measurement pass (one 8-GPU box): sharded verification at N=1,2,4,8, all-pairs search at N=1,8,
# batched registration at N=8.  Every line is the bench's own JSON (device time, max over ranks).
set -u
out=gpurun_out/scale
mkdir -p $out
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@"; }
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --workload verify --steps 2 > $out/verify_$n.json 2> $out/verify_$n.err
  else tr $n --workload verify --steps 2 > $out/verify_$n.json 2> $out/verify_$n.err; fi
done
python bench.py --workload allpairs --unique 1024 --steps 2 > $out/allpairs_1.json 2> $out/allpairs_1.err
tr 8 --workload allpairs --unique 1024 --steps 2 > $out/allpairs_8.json 2> $out/allpairs_8.err
tr 8 --steps 20 > $out/align_8.json 2> $out/align_8.err
nvidia-smi -L > $out/gpus.txt
grep -h "NVLS\|NVLink" $out/*.err | head -5 >> $out/gpus.txt
tail -n 3 $out/*.json
