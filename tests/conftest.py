import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
FILES = ("dict", "occ", "parse", "last", "sai")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


class Golden:
    def __init__(self):
        with open(os.path.join(GOLDEN_DIR, "cases.json")) as f:
            self.meta = json.load(f)["cases"]
        self.npz = np.load(os.path.join(GOLDEN_DIR, "golden.npz"))

    def names(self):
        return [c["name"] for c in self.meta]

    def case(self, name):
        c = dict(next(m for m in self.meta if m["name"] == name))
        c["input"] = self.npz[name + "/input"].tobytes()
        for ext in FILES:
            c[ext] = self.npz[f"{name}/{ext}"].tobytes()
        return c


_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden


def golden_names():
    with open(os.path.join(GOLDEN_DIR, "cases.json")) as f:
        return [c["name"] for c in json.load(f)["cases"]]


def golden_dicz():
    """name -> (case meta, .dicz bytes written by newscanNT.x -c); tools/make_golden_dicz.py"""
    with open(os.path.join(GOLDEN_DIR, "cases_dicz.json")) as f:
        meta = json.load(f)["cases"]
    npz = np.load(os.path.join(GOLDEN_DIR, "golden_dicz.npz"))
    return {c["name"]: (c, npz[c["name"] + "/dicz"].tobytes()) for c in meta}


def assert_same_files(got, want, what=""):
    """Byte-compare the five outputs; `got`/`want` expose .dict .occ .parse .last .sai or keys."""
    for ext in FILES:
        g = getattr(got, ext) if not isinstance(got, dict) else got[ext]
        w = getattr(want, ext) if not isinstance(want, dict) else want[ext]
        g, w = bytes(g), bytes(w)
        if g != w:
            n = min(len(g), len(w))
            first = next((i for i in range(n) if g[i] != w[i]), n)
            raise AssertionError(f"{what}: .{ext} differs (len {len(g)} vs {len(w)}, first diff at {first})")
