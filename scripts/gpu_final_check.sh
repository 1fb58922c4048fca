#!/bin/bash
# Short validation call (little box time left): the tests added last first, then as much of the whole GPU suite
# as fits, then smoke.  Every stage has its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -m pytest -m gpu -q -x --timeout 150 \
  tests/test_gpu_rules.py::test_perft_rule_corner_positions tests/test_gpu_rules.py::test_movegen_most_legal_moves \
  tests/test_gpu_rules.py::test_rules_on_unreachable_random_positions \
  tests/test_gpu_tree.py::test_search_from_unreachable_roots_matches_oracle \
  tests/test_gpu_api.py::test_c_abi_argument_errors > gpurun_out/pytest_new.log 2>&1
echo "== new tests: exit $? at $((SECONDS-T0)) s"; tail -15 gpurun_out/pytest_new.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke: exit $? at $((SECONDS-T0)) s"; tail -2 gpurun_out/smoke.log
timeout ${FULL_TIMEOUT:-600} python -m pytest tests -m gpu -x --timeout 300 -v --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu: exit $? at $((SECONDS-T0)) s"; tail -25 gpurun_out/pytest_gpu.log
