# one ncu --set full capture of score_kernel (bench's scoring pass), brought back in gpurun_out/
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 4 -c 1 -f -o gpurun_out/prof_score \
    python bench.py --steps 3 --warmup 3 --verify-candidates 0 --sustain 0 --no-cpu-baseline > gpurun_out/prof_score.log 2>&1
tail -3 gpurun_out/prof_score.log
