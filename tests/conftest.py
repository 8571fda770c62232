import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def rel(a, b):
    """||a-b||_2 / ||b||_2 for arrays, |a-b|/|b| for scalars (SURVEY.md 7.1 parity metric)."""
    a = np.asarray(a); b = np.asarray(b)
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = a.astype(dt); b = b.astype(dt)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    return num / den if den > 0 else num


@pytest.fixture(scope="session")
def cman():
    """image/cman.png of the reference (256x256 uint8), stored as a fixture."""
    return np.load(os.path.join(GOLDEN, "cman_u8.npy")).astype(np.float64)


@pytest.fixture(scope="session")
def boat():
    """images/boat.png of the reference (512x512 uint8), stored as a fixture."""
    return np.load(os.path.join(GOLDEN, "boat_u8.npy")).astype(np.float64)


def kat_image(n=256):
    """Formula-defined image of SURVEY.md 3.3 (early-stop KAT)."""
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return 127.5 + 100 * np.sin(2 * np.pi * 3 * i / n) * np.cos(2 * np.pi * 5 * j / n) + 0.37 * ((7 * i + 13 * j) % 11)
