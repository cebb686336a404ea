mkdir -p gpurun_out
for t in score multi preprocess mapping; do
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200 python -m pytest tests/test_gpu_$t.py -m gpu -q -x -k "not full_size" > gpurun_out/racecheck_$t.log 2>&1
  echo "$t: $(grep -c 'Race reported' gpurun_out/racecheck_$t.log) reports; $(tail -1 gpurun_out/racecheck_$t.log)"
done
