"""Test configuration: paths, the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` covers the oracle against the golden vectors, the host logic and the C-ABI
surface (no compute calls); ``-m gpu`` holds the parity tests proper, which call the CUDA kernels
through the C ABI and compare with the oracle / the committed goldens.  Nothing here reads
/root/reference.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio-adaptive-tokenizer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self):
        gdir = os.path.join(ROOT, "tests", "golden")
        self.arrays = np.load(os.path.join(gdir, "golden_v1.npz"))
        with open(os.path.join(gdir, "MANIFEST.json")) as f:
            self.manifest = json.load(f)
        self.cases = self.manifest["cases"]

    def get(self, case, name):
        return self.arrays[f"{case}/{name}"]

    def has(self, case, name):
        return f"{case}/{name}" in self.arrays.files

    def wave(self, case):
        """Regenerate the case's input from its recorded recipe (deterministic numpy generators)."""
        from aat_b200 import synth

        if self.has(case, "wave"):
            return self.get(case, "wave")
        gen = self.cases[case]["gen"]
        env = {"bursty_speech": synth.bursty_speech, "stationary_noise": synth.stationary_noise,
               "zeros": lambda n: np.zeros(n), "znorm": synth.znorm, "float64": lambda a: a.astype(np.float64)}
        if gen.startswith("zeros("):
            return np.zeros(int(gen[6:gen.index(")")]), dtype=np.float64)
        return eval(gen, {"__builtins__": {}}, env)


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co

    co.build()
    return co
