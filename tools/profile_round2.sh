# Round-2 evidence run (one GPU), in parts small enough for gpurun_out/ (64 MiB per call):
#   bash tools/profile_round2.sh bench | ncu1 | ncu2 | ncu3
# tools/ncu_summary.py turns the .ncu-rep files into profiles/r02/*.md here.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
case "$1" in
bench)
  python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
  for w in track multi; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2>/dev/null; done
  python bench.py --workload track --voxel 0 --no-cpu-baseline > gpurun_out/r02_bench_track_novoxel.json 2>/dev/null
  python bench.py --beams 721 --no-cpu-baseline --verify-candidates 0 --sustain 0 > gpurun_out/r02_bench_721.json 2>/dev/null
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"icp_|score_|preprocess|clip|scan_|best_of|project|correspond|merge|beam" \
      -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --verify-candidates 4096 --sustain 0 > gpurun_out/launch_bench.log 2>&1
  cat gpurun_out/r02_bench.json ;;
ncu1)
  $NCU -k regex:icp_fused2 -s 4 -c 1 -o gpurun_out/prof_icp2 python bench.py --steps 3 --warmup 3 --verify-candidates 0 --sustain 0 --no-cpu-baseline > /dev/null 2>&1
  $NCU -k regex:score_kernel -s 4 -c 1 -o gpurun_out/prof_score python bench.py --steps 3 --warmup 3 --verify-candidates 0 --sustain 0 --no-cpu-baseline > /dev/null 2>&1
  $NCU -k regex:icp_multi2 -s 3 -c 1 -o gpurun_out/prof_multi2 python bench.py --workload multi --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 ;;
ncu2)
  $NCU -k regex:"preprocess_kernel|clip_kernel|scan_pack" -s 12 -c 3 -o gpurun_out/prof_track python bench.py --workload track --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 ;;
ncu3)
  ncu --set full --clock-control none -f -k regex:"icp_general|icp_stream|clip_voxel|merge_kernel|project_kernel|correspond_kernel" -c 12 \
      -o gpurun_out/prof_service python tools/exercise_kernels.py > gpurun_out/exercise.log 2>&1
  tail -2 gpurun_out/exercise.log ;;
esac
ls -la gpurun_out | tail -12
