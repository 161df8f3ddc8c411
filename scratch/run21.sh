export PYTHONUNBUFFERED=1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pass_first_vs_reference or pass_next_function_boundary or plan_chain or degenerate or single_window or huge_and or windows_unaligned" 2>&1 | tail -6
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pass_next_function_boundary or degenerate or single_window" 2>&1 | tail -6
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pass_next_function_boundary or degenerate" 2>&1 | tail -6
