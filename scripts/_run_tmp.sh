export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_minimizers.py tests/test_gpu_compact.py tests/test_gpu_packed.py -x -q -m gpu 2>&1 | tail -6
echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "geometry or ragged or wide or straddl or short" 2>&1 | tail -6
echo "racecheck rc=$?"
