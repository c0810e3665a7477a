"""The drop-in surface cannot drift unnoticed: every member of the ORB-SLAM3 / MS-SLAM classes that the sparsifier path
touches is declared by the shims (ms_slam_b200/host/SlamShims.h) and by the replacement class
(ms_slam_b200/host/MapSparsification.h) with the signature the reference's own headers give it.  The reference's signatures
are committed as tests/golden/reference_signatures.json (made by tests/golden/make_reference_signatures.py); where
/root/reference is present the fixture itself is checked against the live headers."""
import importlib.util
import json
import os

import pytest

from conftest import ROOT, GOLDEN

spec = importlib.util.spec_from_file_location("make_reference_signatures", os.path.join(GOLDEN, "make_reference_signatures.py"))
mrs = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mrs)

HOST = os.path.join(ROOT, "ms_slam_b200", "host")
SHIM_FILES = {"MapPoint": "SlamShims.h", "KeyFrame": "SlamShims.h", "Map": "SlamShims.h", "Atlas": "SlamShims.h", "LoopClosing": "SlamShims.h",
              "MapSparsification": "MapSparsification.h"}
# deliberate, documented deviations of the shims (SlamShims.h says why)
ALLOWED = {("KeyFrame", "mnMapSaprsificationId*"), ("KeyFrame", "N*")}


def test_shims_declare_the_reference_surface():
    ref = json.load(open(os.path.join(GOLDEN, "reference_signatures.json")))
    ours = mrs.extract(HOST, SHIM_FILES)
    problems = []
    for cls, members in ref.items():
        for name, want in members.items():
            got = ours[cls].get(name, [])
            if not got:
                problems.append(f"{cls}::{name} is not declared by the host mirror")
            elif not set(got) & set(want) and (cls, name) not in ALLOWED:
                problems.append(f"{cls}::{name}: ours {got} vs reference {want}")
    assert not problems, "\n".join(problems)


def test_fixture_matches_the_live_reference_headers():
    ref_dir = os.path.join(mrs.REF, "include")
    if not os.path.isdir(ref_dir):
        pytest.skip("the reference tree is not on this box; the committed fixture stands in for it")
    live = mrs.extract(ref_dir)
    assert live == json.load(open(os.path.join(GOLDEN, "reference_signatures.json")))
