import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


REFERENCE_ROOT = "/root/reference"
HAVE_REFERENCE = os.path.isdir(os.path.join(REFERENCE_ROOT, "general_motion_retargeting"))
needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason="reference checkout not present (GPU box)")

ALL_PAIRS = [
    ("smplx", "unitree_g1"), ("smplx", "booster_t1"), ("smplx", "stanford_toddy"), ("smplx", "fourier_n1"),
    ("smplx", "engineai_pm01"), ("smplx", "kuavo_s45"), ("smplx", "hightorque_hi"),
    ("bvh", "unitree_g1"), ("bvh", "booster_t1"), ("bvh", "booster_t1_4dof"), ("bvh", "fourier_n1"),
    ("bvh", "stanford_toddy"), ("bvh", "engineai_pm01"), ("fbx", "unitree_g1"),
]


@pytest.fixture(scope="session")
def built():
    """Native pieces the CPU tests need (oracle .so, emulator .so); the CUDA lib is built by
    __graft_entry__.build() and only dlopen'ed here."""
    import __graft_entry__ as g
    g.build()
    return True
