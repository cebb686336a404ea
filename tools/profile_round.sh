set -x
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>> gpurun_out/r01_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"icp_|preprocess|clip|scan_|best_of|project|correspond|merge" -c 40 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or zbuffer_ties or ragged or status_codes or sensor_offset" > gpurun_out/r01_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or zbuffer_ties or ragged or status_codes or sensor_offset or score_batch or zero_pairs" > gpurun_out/r01_memcheck.log 2>&1
tail -n 4 gpurun_out/r01_racecheck.log gpurun_out/r01_memcheck.log
for w in track multi; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r01_bench_$w.json 2>/dev/null; done
python bench.py --workload track --voxel 0 --no-cpu-baseline > gpurun_out/r01_bench_track_novoxel.json 2>/dev/null
python bench.py --workload verify --candidates 16384 --steps 5 --no-cpu-baseline > gpurun_out/r01_bench_verify16k.json 2>/dev/null
cat gpurun_out/r01_bench.json
