#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
# attempt compute-sanitizer memcheck on a small slice of the parity suite
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $O/memcheck.log python -m pytest tests/test_search_step_gpu.py tests/test_tree_gpu.py tests/test_env_gpu.py -m gpu -q -x -k "not 4096 and not 2048 and not pipeline" > $O/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -n 5 $O/memcheck_pytest.log
tail -n 25 $O/memcheck.log
