import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


# Make the real reference importable (when it is mounted) BEFORE pints_b200 is
# imported, so the mirror classes pick up the reference's base classes.  Set
# PINTS_B200_NO_PINTS=1 to run the suite as on the GPU box (no reference).
# The unmodified reference is installed into the git-ignored baseline/_ref by
# __graft_entry__.build() and travels to the GPU box with the snapshot;
# /root/reference (build container only) is the second choice.
_REF = os.environ.get('PINTS_REFERENCE')
if _REF is None:
    _REF = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(_REF, 'pints')):
        _REF = '/root/reference'
if os.environ.get('PINTS_B200_NO_PINTS') != '1' and \
        os.path.isdir(os.path.join(_REF, 'pints')) and _REF not in sys.path:
    sys.path.append(_REF)
    sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run with -m gpu on a B200)')
    config.addinivalue_line(
        'markers', 'reference: needs the real reference (baseline/_ref)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


@pytest.fixture(scope='session')
def golden():
    return load_golden


def reference_pints():
    """The real reference, when it is mounted (build container only)."""
    if os.environ.get('PINTS_B200_NO_PINTS') == '1':
        return None
    try:
        import pints
        import pints.toy  # noqa: F401
    except Exception:
        return None
    return pints


@pytest.fixture(scope='session')
def pints_ref():
    p = reference_pints()
    if p is None:
        pytest.skip('reference not available')
    return p
