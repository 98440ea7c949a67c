"""The C ABI as its three bindings see it (no GPU, no Fortran compiler needed).

* include/cosmob200.h is compiled with gcc and `_Static_assert`ed against the layout numbers that
  cosmomc_b200/fortran/Calculator_B200.f90 states for its `bind(C)` mirror of cb200_config;
* the field list of the Fortran `type, bind(C) :: cb200_config` is parsed and compared, name by name and kind by kind,
  with the header's struct (a bind(C) derived type with the same component order and interoperable kinds has the C
  layout: Fortran 2008 15.3.4);
* the ctypes mirror (cosmomc_b200/lib.py) must have the same size and offsets;
* every entry point the header declares is exported by libcosmob200.so, and cb200_create refuses a stale mirror.
"""
import ctypes as C
import os
import re
import subprocess
import tempfile

import helpers as H

HDR = os.path.join(H.ROOT, "include", "cosmob200.h")
F90 = os.path.join(H.ROOT, "cosmomc_b200", "fortran", "Calculator_B200.f90")


def header_fields():
    src = open(HDR).read()
    body = re.search(r"typedef struct cb200_config \{(.*?)\} cb200_config;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        kind, names = decl.split(None, 1)
        out += [(n.strip(), kind) for n in names.split(",")]
    return out


def fortran_layout():
    txt = open(F90).read()
    nums = {}
    for line in txt.splitlines():
        if "ABI-LAYOUT" in line:
            for k, v in re.findall(r"(\w+)=(\d+)", line.split("ABI-LAYOUT", 1)[1]):
                nums[k] = int(v)
    body = re.search(r"type, bind\(C\) :: cb200_config\n(.*?)end type cb200_config", txt, re.S).group(1)
    fields = []
    for line in body.splitlines():
        m = re.match(r"\s*(integer\(c_int\)|real\(c_double\)) :: (.*)", line)
        if m:
            kind = "int" if "c_int" in m.group(1) else "double"
            fields += [(n.strip(), kind) for n in m.group(2).split(",")]
    return nums, fields


def test_fortran_mirror_matches_header_fields():
    nums, ffields = fortran_layout()
    assert ffields == header_fields()


def test_header_layout_static_asserts():
    nums, _ = fortran_layout()
    lines = ['#include <stddef.h>', '#include "cosmob200.h"',
             '_Static_assert(sizeof(cb200_config) == %d, "sizeof");' % nums.pop("sizeof")]
    assert {n for n, _ in header_fields()} == set(nums), "the Fortran comment must state every field's offset"
    for name, off in nums.items():
        lines.append('_Static_assert(offsetof(cb200_config, %s) == %d, "%s");' % (name, off, name))
    lines.append("int main(void) { return 0; }")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "abi.c")
        open(c, "w").write("\n".join(lines) + "\n")
        r = subprocess.run(["gcc", "-std=c11", "-I", os.path.dirname(HDR), "-c", c, "-o", os.path.join(d, "abi.o")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_ctypes_mirror_and_exports():
    from cosmomc_b200 import lib
    nums, _ = fortran_layout()
    assert C.sizeof(lib.Config) == nums["sizeof"]
    for name, _ in header_fields():
        assert getattr(lib.Config, name).offset == nums[name], name
    L = lib.load()
    assert L.cb200_config_size() == nums["sizeof"]
    declared = set(re.findall(r"\b(cb200_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)))
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(lib.EXPORTS) <= declared


def test_create_refuses_a_stale_mirror():
    from cosmomc_b200 import lib
    L = lib.load()
    cfg = lib.Config()
    L.cb200_default_config(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(lib.Config)
    cfg.struct_size -= 8  # what an out-of-date binding would pass
    h = C.c_void_p()
    assert L.cb200_create(C.byref(cfg), C.byref(h)) == -3 and not h


def _header_protos():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|void|const char\*)\s+(cb200_\w+)\s*\(([^;{]*?)\)\s*;", src, re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_fortran_bindings_name_exported_symbols_with_matching_arity():
    """Every `bind(C, name='cb200_...')` interface of the Fortran glue names an entry point the header declares, with
    the same number of arguments (the Fortran side cannot be compiled here; this catches a binding that drifted)."""
    protos = _header_protos()
    fdir = os.path.join(H.ROOT, "cosmomc_b200", "fortran")
    seen = set()
    for fn in sorted(os.listdir(fdir)):
        if not fn.endswith(".f90"):
            continue
        txt = open(os.path.join(fdir, fn)).read().replace("&\n", " ")
        for m in re.finditer(r"(?:function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(cb200_\w+)'\)", txt):
            fname, args, cname = m.group(1), m.group(2), m.group(3)
            nargs = len([a for a in args.split(",") if a.strip()])
            assert cname in protos, (fn, cname)
            assert nargs == protos[cname], (fn, cname, nargs, protos[cname])
            seen.add(cname)
    assert {"cb200_create", "cb200_upload_sources_packed", "cb200_powers", "cb200_like_add_cmblikes",
            "cb200_like_add_pliklite", "cb200_loglike_batch"} <= seen


def test_camb_patch_is_a_well_formed_unified_diff():
    """cosmomc_b200/fortran/camb_sources_only.patch: the CAMB-side hook shipped as a diff (applies with `patch -p1` in the
    reference root; checked against the reference checkout when it is present, i.e. in the build container)."""
    p = os.path.join(H.ROOT, "cosmomc_b200", "fortran", "camb_sources_only.patch")
    txt = open(p).read()
    assert txt.startswith("--- a/camb/cmbmain.f90") and "+++ b/camb/cmbmain.f90" in txt
    assert "cmbmain_sources_only" in txt and txt.count("\n@@ ") == 2
    ref = "/root/reference"
    if os.path.isdir(os.path.join(ref, "camb")):
        import shutil
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "camb"))
            shutil.copy(os.path.join(ref, "camb", "cmbmain.f90"), os.path.join(d, "camb", "cmbmain.f90"))
            r = subprocess.run(["patch", "-p1", "--dry-run", "-i", p], cwd=d, capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
