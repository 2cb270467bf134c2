import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """import vido-slam_b200/ (hyphenated directory) as module vido_slam_b200"""
    if "vido_slam_b200" in sys.modules:
        return sys.modules["vido_slam_b200"]
    path = os.path.join(ROOT, "vido-slam_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("vido_slam_b200", path,
                                                  submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vido_slam_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()
