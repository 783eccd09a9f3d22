import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Native pieces are built artefacts (git-ignored): build what is missing so a fresh checkout is testable.
    lib = os.path.join(ROOT, "centroidalplanner_b200", "libcplb.so")
    orc = os.path.join(ROOT, "oracle", "libcpl_oracle.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "centroidalplanner_b200", "csrc"), "-j", "8"],
                              stdout=subprocess.DEVNULL)
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libcpl_oracle.so"], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def pytest_terminal_summary(terminalreporter):
    """How often the expanded-square allowance (tests/helpers.py) was actually needed in this session."""
    try:
        import helpers
    except Exception:
        return
    a = helpers.ALLOWANCE
    if a["pow_entries"]:
        terminalreporter.write_line(f"parity: {a['pow_entries']} pow-downstream entries compared, {a['entries']} beyond the plain 1e-12 bar "
                                    f"(passed through the Q5 allowance, factor {helpers.Q5_FACTOR}); worst entry {a['worst_vs_plain_bar']:.3g} x the plain bar")
