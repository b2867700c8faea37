"""
TEST INFRASTRUCTURE -- container only.  Runs the reference's own fast tests
(raoteh/sampler/tests) through the compatibility shim, to gate the shim and
the pyfelscore stand-in before they are trusted to generate golden vectors.
Expected: 37 passed, 1 failed
(test_tmjp.py::test_primary_trajectory_log_likelihood is a pre-existing
non-identity in the reference) with the long print-only tests deselected.

usage: python oracle/run_reference_tests.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

ref_shim.load_reference()
import pytest  # noqa: E402

tests = os.path.join(ref_shim.REFERENCE_ROOT, 'raoteh', 'sampler', 'tests')
deselect = [
    # print-only Monte-Carlo tables that end in `raise Exception`
    'test_sampler.py::TestRaoTehSampler::test_gen_histories_primary_process_entropy',
    'test_sample_mjp.py',
    'test_sample_tmjp.py',
]
args = ['-p', 'no:cacheprovider', '--rootdir', '/tmp', '-rN', '--tb=line',
        '-k', 'not entropy and not differential and not slow', tests]
for d in deselect[1:]:
    args += ['--ignore', os.path.join(tests, d)]
sys.exit(pytest.main(args))
