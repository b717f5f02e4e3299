import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_terminal_summary(terminalreporter):
    """Fast-kernel contract tally (tests/gpu_helpers.py): values compared against the 1e-9 bound of north_star and how
    many needed the conditioning slack (all of them ill conditioned by construction of the assertion)."""
    gh = sys.modules.get("gpu_helpers")
    if gh is not None and gh.SLACK["values"]:
        s = gh.SLACK
        terminalreporter.write_line(
            "fast-kernel contract: %d values compared, max |delta| %.3g, %d needed more than 1e-9 "
            "(largest condition number among them %.3g), %d NaN/inf pattern differences on ill-conditioned values"
            % (s["values"], s["max_err"], s["needed_slack"], s["max_kappa_of_slack"], s["nan_pattern_forgiven"]))
