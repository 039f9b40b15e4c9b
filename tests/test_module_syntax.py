"""The Voreen module (custommodules/b200raycast) must compile against the reference's own headers.

A syntax-only pass (g++ -fsyntax-only, no Qt / GL libraries needed) is what this container can do; it catches wrong include
paths, const errors and API drift.  Skipped where /root/reference is absent (the GPU box)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = "/root/reference/voreen"


@pytest.mark.skipif(not os.path.isdir(os.path.join(V, "include", "voreen")), reason="reference headers not present")
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_voreen_module_sources_pass_the_syntax_target():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "custommodules", "b200raycast"), "syntax", f"V={V}"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    for f in ("b200raycastmodule.cpp", "b200singlevolumeraycaster.cpp", "b200multivolumeraycaster.cpp", "b200simpleraycaster.cpp", "b200rgbraycaster.cpp"):
        assert f in r.stdout


def test_basemodule_guard_patch_applies():
    patch = os.path.join(ROOT, "custommodules", "b200raycast", "patches", "basemodule_guard.patch")
    target = os.path.join(V, "modules", "base", "basemodule.cpp")
    if not os.path.exists(target) or shutil.which("patch") is None:
        pytest.skip("reference tree or patch(1) not present")
    r = subprocess.run(["patch", "--dry-run", "-p1", "-d", V, "-i", patch, "-o", "/dev/null"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
