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


def test_r_scripts_draw_the_reference_plots_from_what_the_library_returns():
    """NanoTelGPU.R calls the reference's own plot functions with the window tables of the library (not a comment),
    and plot_density_vectors.R reads exactly the columns the R-free driver writes to density_vectors/read<Serial>.csv."""
    import re
    rdir = os.path.join(ROOT, "telomere-analyzer_b200", "R")
    gpu = open(os.path.join(rdir, "NanoTelGPU.R")).read()
    code = "\n".join(l.split("#")[0] for l in gpu.splitlines())
    for fn in ("plot_single_telo_with_gray_area", "plot_single_telo_with_tvr", "writeXStringSet", "ntl_R_scan_xstringset"):
        assert fn in code, fn
    assert code.count("do.call(plot_single_telo_with_gray_area") == 3 and code.count("do.call(plot_single_telo_with_tvr") == 3
    plot = open(os.path.join(rdir, "plot_density_vectors.R")).read()
    py = open(os.path.join(ROOT, "telomere-analyzer_b200", "nanotel_b200", "nanotel.py")).read()
    # the per-read files are written by the library (csrc/ntl_writer.cpp, called from nanotel.py)
    wr = open(os.path.join(ROOT, "telomere-analyzer_b200", "csrc", "ntl_writer.cpp")).read()
    assert "ntl_write_read_outputs" in py and "/density_vectors" in wr and "density_vectors" in plot
    assert '"ID,start_index,end_index"' in wr and '",density"' in wr and '",class"' in wr
    for col in ("ID", "start_index", "end_index", "density", "class"):
        assert col in plot
    for sfx in ("_mismatch", "_mismatch_tvr"):
        assert '"%s"' % sfx in wr and '"%s"' % sfx in plot
    # the summary columns the script reads are the ones write_summary_csv emits
    for col in re.findall(r"row\$(\w+)", plot):
        assert col in py or col in ("Serial",), col
