"""Committed golden results (tests/golden/scenario_plans.json, written by tests/golden/make_golden.py from the pinned
oracle): the oracle must still produce them, the device's row logic compiled for the host must match them, and on a GPU
the CUDA path is compared with the committed bytes directly.  (Last file of the suite on purpose.)"""
import json
import os
import sys

import pytest

import emu
import orc
import scenarios as sc

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402

with open(os.path.join(HERE, "golden", "scenario_plans.json")) as f:
    GOLDEN = json.load(f)
CASES = G.cases()


def _check(run, names=None):
    region = G.region()
    n = 0
    for name, plan in CASES:
        if names is not None and name not in names:
            continue
        want = GOLDEN[name]
        res = run(plan, sc.split_ranges(), region)
        assert res.status == want["status"], (name, res.status, want["status"])
        assert G.canon(res.rows(), name) == want["rows"], name
        n += 1
    return n


def test_golden_file_covers_every_case():
    assert sorted(GOLDEN) == sorted(n for n, _ in CASES) and len(GOLDEN) >= 50


def test_oracle_matches_committed_golden():
    assert _check(orc.dag_handle) == len(CASES)


def test_device_logic_matches_committed_golden():
    assert _check(emu.dag_handle) == len(CASES)


@pytest.mark.gpu
def test_cuda_path_matches_committed_golden():
    from tikv_b200.executor import DagHandler, DeviceRegion
    assert _check(lambda plan, ranges, region: DagHandler(plan, ranges, DeviceRegion(region)).handle_request()) == len(CASES)
