# compute-sanitizer over the GPU suite (memcheck: everything but the full-size cases; racecheck: the kernels with
# hand-rolled shared-memory protocols).  Logs under gpurun_out/ -> copy into profiles/rNN/.
mkdir -p gpurun_out
( echo 'compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "not full_size and not allpairs and not nccl"'
  timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "not full_size and not allpairs and not nccl" 2>&1 | grep -v "^$" | tail -40 ) > gpurun_out/memcheck.log
( echo 'compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_score.py tests/test_gpu_multi.py tests/test_gpu_preprocess.py tests/test_gpu_mapping.py -m gpu -q -x -k "not full_size"'
  timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_score.py tests/test_gpu_multi.py tests/test_gpu_preprocess.py tests/test_gpu_mapping.py -m gpu -q -x -k "not full_size" 2>&1 | grep -v "^$" | tail -40 ) > gpurun_out/racecheck.log
tail -5 gpurun_out/memcheck.log gpurun_out/racecheck.log
