# Sanitizer / generic-path pass on the final library (one B200).  Output: gpurun_out/sanitizer.txt
export PATH=/usr/local/cuda/bin:$PATH
out=gpurun_out/sanitizer.txt
{
echo "# 1. the whole GPU suite on the GENERIC scatter path"
echo "FSGPU_FORCE_GENERIC=1 python -m pytest tests -m gpu -x -q"
FSGPU_FORCE_GENERIC=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
echo "# 2. compute-sanitizer --tool memcheck --error-exitcode 9"
K2="element_stiffness or thickness or deterministic or triangle or pageable or device_csys or coo_to_csc or q4rscomp_per_point or assembled_stiffness or permuted or beam_operators"
echo "python -m pytest tests/test_gpu_parity2.py tests/test_gpu_parity.py -x -q -k \"$K2\""
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_parity.py -x -q -k "$K2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tail -3
echo "# 3. compute-sanitizer --tool racecheck --error-exitcode 9"
K3="element_stiffness or deterministic_gather or assembled_composite"
echo "python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -x -q -k \"$K3\""
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -x -q -k "$K3" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY" | tail -3
} > $out 2>&1
cat $out
