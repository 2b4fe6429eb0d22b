"""pytest plugin (``-p b200_enable_plugin``): run somebody else's tests with the B200 backend
enabled.  Used by tests/test_gpu_reference_suite.py to rerun the REFERENCE's own test files."""
import json
import os


def pytest_configure(config):
    import gstools_b200

    gstools_b200.enable()


def pytest_unconfigure(config):
    import gstools_b200

    out = os.environ.get("GSB200_SUITE_REPORT")
    if out:
        names = ("launches", "direct_calls", "separable_calls", "krige_calls")
        json.dump({n: gstools_b200.get_counter(n) for n in names}, open(out, "w"))
    gstools_b200.disable()
