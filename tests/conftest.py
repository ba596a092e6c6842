import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def repo_root():
    return REPO


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Worst measured parity error per label (tests/helpers.check_err), next to its tolerance."""
    import json
    from tests import helpers
    if not helpers.ERRLOG:
        return
    worst = {}
    for label, err, tol in helpers.ERRLOG:
        if label not in worst or err > worst[label][0]:
            worst[label] = (err, tol)
    terminalreporter.write_sep('-', 'measured parity errors (worst per label) vs tolerance')
    for label in sorted(worst):
        err, tol = worst[label]
        terminalreporter.write_line(f'{label:<58s} {err:10.3e}   tol {tol:8.1e}   margin x{tol / max(err, 1e-30):.1f}')
    try:
        out = os.path.join(REPO, 'gpurun_out')
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'parity_errors.json'), 'w') as f:
            json.dump({k: {'err': v[0], 'tol': v[1]} for k, v in worst.items()}, f, indent=1, sort_keys=True)
    except OSError:
        pass
