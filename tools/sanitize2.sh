# compute-sanitizer synccheck and initcheck over the kernels with hand-rolled protocols
mkdir -p gpurun_out
for tool in synccheck initcheck; do
  ( echo "compute-sanitizer --tool $tool python -m pytest tests/test_gpu_score.py tests/test_gpu_multi.py tests/test_gpu_preprocess.py tests/test_gpu_mapping.py tests/test_gpu_options.py -m gpu -q -x -k 'not full_size and not deterministic'"
    timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_score.py tests/test_gpu_multi.py tests/test_gpu_preprocess.py tests/test_gpu_mapping.py tests/test_gpu_options.py -m gpu -q -x -k "not full_size and not deterministic" 2>&1 | grep -v "^$" | tail -25 ) > gpurun_out/$tool.log
  tail -3 gpurun_out/$tool.log
done
