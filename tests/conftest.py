import json
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionfinish(session, exitstatus):
    """Write the errors the parity tests measured (tests/util.py: record) to one JSON file."""
    try:
        import util
    except Exception:
        return
    if not util.PARITY:
        return
    out = os.environ.get("HARP_PARITY_OUT") or os.path.join(ROOT, "gpurun_out", "parity.json")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        prev = {}
        if os.path.exists(out) and os.environ.get("HARP_PARITY_APPEND"):
            prev = json.load(open(out))
        prev.update(util.PARITY)
        with open(out, "w") as f:
            json.dump(prev, f, indent=1, sort_keys=True)
    except Exception as e:  # never fail the run for the log
        print(f"[parity log] not written: {e}")
