#!/bin/bash
# the driver's scaling pass, run by hand on one 8-GPU box: the default bench line (which carries the sharded
# verification through ls2d_verify_sharded_nccl) at N = 8 and N = 4; N = 1, 2 are in profiles/r02/bench_n{1,2}.json
set -u
out=gpurun_out/scale_r02
mkdir -p $out
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 2> $out/bench_n$n.err | tail -1 > $out/bench_n$n.json
done
nvidia-smi -L > $out/gpus.txt
python - <<'PY'
import json
for n in (8, 4):
    d = json.loads(open("gpurun_out/scale_r02/bench_n%d.json" % n).read())
    print(n, d["value"], d["e2e"]["value"], d["verify"]["value"], d["verify"]["ms_per_step"], d["verify"]["winner"])
PY
