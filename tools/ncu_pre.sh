# one ncu --set full capture of preprocess_kernel (the tracker's frame step), brought back in gpurun_out/
ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 4 -c 1 -f -o gpurun_out/prof_pre \
    python bench.py --workload track --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof_pre.log 2>&1
tail -3 gpurun_out/prof_pre.log
