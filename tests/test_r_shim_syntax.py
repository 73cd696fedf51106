"""CPU: the .Call glue for an R host (telomere-analyzer_b200/R/r_shim.c) cannot be built here -- R is not in the image
-- but it can be parsed and type-checked: gcc -fsyntax-only against tests/r_stub/ (prototypes of the R C API written
from "Writing R Extensions") and the real include/nanotel_b200.h.  Catches a C ABI that drifted away from the glue."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not found")
def test_r_glue_type_checks_against_the_c_abi():
    cmd = ["gcc", "-fsyntax-only", "-std=c11", "-Wall", "-Wextra", "-Wno-cast-function-type",
           "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types", "-Werror=int-conversion",
           "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "telomere-analyzer_b200", "R", "r_shim.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0 and "warning" not in r.stderr, r.stderr


def test_r_glue_registers_what_the_r_script_calls():
    src = open(os.path.join(ROOT, "telomere-analyzer_b200", "R", "r_shim.c")).read()
    rsrc = open(os.path.join(ROOT, "telomere-analyzer_b200", "R", "NanoTelGPU.R")).read()
    import re
    registered = set(re.findall(r'\{"(ntl_R_\w+)"', src))
    called = set(re.findall(r'\.Call\("?(ntl_R_\w+)"?', rsrc))
    assert registered and called and called <= registered, (registered, called)
