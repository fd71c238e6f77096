"""Runs the bf16 fixture tests of tests/test_msda_gpu.py 40 times in one process (reduction order differs from run to run:
a bound that holds once may not hold always).  `python tools/loop_bf16_fixture_tests.py` on the GPU box; prints `loop ok 120`."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_msda_gpu as t
n = 0
for i in range(40):
    for name in ("msda_tiny_U", "msda_pyr4_S", "msda_pyr5_D"):
        t.test_bf16_value_path_matches_reference_fixtures(name); n += 1
print("loop ok", n)
