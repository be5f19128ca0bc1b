"""CPU tier: the C++ drop-in (host/ORBextractor_b200.cc) compiles against the REFERENCE'S OWN header
include/ORBextractor.h (signatures unchanged).  Needs /root/reference, so it only runs in the build container."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/include"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ORBextractor.h")), reason="reference checkout not present")
def test_shim_compiles_against_reference_header(tmp_path):
    host = os.path.join(ROOT, "orb_slam3_detailed_comments_b200", "host")
    obj = str(tmp_path / "shim.o")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-c", os.path.join(host, "ORBextractor_b200.cc"),
                           "-I", os.path.join(host, "cvshim"), "-I", REF, "-I", os.path.join(ROOT, "include"), "-o", obj])
    syms = subprocess.check_output(["nm", "-C", obj], text=True)
    assert "ORB_SLAM3::ORBextractor::ORBextractor(int, float, int, int, int)" in syms
    assert "ORB_SLAM3::ORBextractor::operator()(" in syms
    for used in ("orbx_create", "orbx_extract", "orbx_get_tables", "orbx_download_level"):
        assert f"U {used}" in syms
