"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/escort_b200.h
declares; argument validation returns error codes instead of aborting (no compute is launched without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    return C.CDLL(os.path.join(ROOT, "caffe_escoin_b200", "libescort_b200.so"))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "escort_b200.h")).read()
    return sorted(set(re.findall(r"ESCORT_API[^;(]*?\b(escort_\w+)\s*\(", txt)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("escort_pack_csr", "escort_stretch", "escort_copy_input", "escort_sconv_padded", "escort_plan_create",
              "escort_sconv_forward", "escort_sconv_backward_data", "escort_sconv_backward_weight",
              "escort_bias_backward", "escort_refresh_values", "escort_allreduce_grads"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export " + s


def test_python_binding_lists_every_declared_symbol(lib):
    from caffe_escoin_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_bad_arguments_return_codes_not_aborts(lib):
    lib.escort_last_error.restype = C.c_char_p
    assert lib.escort_pack_csr(0, 0, None, None, None, None, None, None, None) == -1
    assert b"escort_pack_csr" in lib.escort_last_error()
    assert lib.escort_plan_create(None, None, None, None, 0, None, None) == -1
    assert lib.escort_sconv_forward(None, 1, None, None, 0, None, None) == -1
    assert lib.escort_plan_destroy(None) == 0
    assert lib.escort_allreduce_grads(None, None, C.c_size_t(0), C.c_float(1.0), None) == 0


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure; nothing under caffe_escoin_b200/ may reference it."""
    pkg = os.path.join(ROOT, "caffe_escoin_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "escort_oracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)


def test_generated_variants_are_consistent():
    """tools/gen_interp.py: one generated handler chain per variant, unique kernel names, and the variant table the
    library was built from lists the same count (a stale generated/ directory would mis-number the variants)."""
    import subprocess
    import sys
    gen = os.path.join(ROOT, "tools", "gen_interp.py")
    n = int(subprocess.run([sys.executable, gen, "--count"], capture_output=True, text=True, check=True).stdout)
    gdir = os.path.join(ROOT, "caffe_escoin_b200", "csrc", "generated")
    lst = open(os.path.join(gdir, "variant_list.inc")).read()
    assert "#define ESCORT_NUM_VARIANTS %d" % n in lst
    names = []
    for i in range(n):
        txt = open(os.path.join(gdir, "interp_v%d.inc" % i)).read()
        m = re.search(r'return "(sconv_tile_\w+)"', txt)
        assert m, i
        names.append(m.group(1))
        assert "template <> struct Interp<%d>" % i in txt
    assert len(set(names)) == n, [x for x in names if names.count(x) > 1]
    # every default the planner names must exist
    src = open(os.path.join(ROOT, "caffe_escoin_b200", "csrc", "sconv_tile.cu")).read()
    for pref in set(re.findall(r'"(sconv_tile_\w+)"', src)):
        assert pref in names, pref


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/escort_b200.h compiles as C99 with -pedantic (what a cgo / ctypes-style binding or
    Caffe's C++ both see), and a C program links against every entry it names."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "escort_b200.h"\n'
                   'int main(void) { escort_geom g; (void)g; return escort_version() == 0; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + inc, "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pkg = os.path.join(ROOT, "caffe_escoin_b200")
    exe = tmp_path / "hdr"
    r = subprocess.run(["gcc", "-std=c99", "-I" + inc, str(src), "-o", str(exe), "-L" + pkg, "-lescort_b200",
                        "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + pkg, "-Wl,-rpath,/usr/local/cuda/lib64"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
