python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fri.py tests/test_gpu_permutation.py -x -q -k "not larger and not 11-20" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_quotient.py -x -q -k "recursion or edge or other" 2>&1 | tail -6
