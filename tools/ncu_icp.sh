# one ncu --set full capture of the headline kernel (bench's align step), brought back in gpurun_out/
ncu --set full --clock-control none --import-source on -k regex:icp_fused2 -s 4 -c 1 -f -o gpurun_out/prof_icp2 \
    python bench.py --steps 3 --warmup 3 --verify-candidates 0 --sustain 0 --no-cpu-baseline > gpurun_out/prof_icp2.log 2>&1
tail -c 300 gpurun_out/prof_icp2.log
