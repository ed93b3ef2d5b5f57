"""The boundary seen from a compiled-language client: tests/c/harness.c (plain C, gcc) against include/hp3d_gpu.h and the raw .so.

* layout: sizeof / offsetof of hp3d_params and hp3d_physics as the C compiler lays them out == the layout ISO_C_BINDING gives the
  `type, bind(C)` declarations of integration/hp3d_gpu_mod.F90 (companion-processor rules: every component at its natural
  alignment, in declaration order) == the ctypes mirror the Python tests use.  Runs without a GPU.
* replay (GPU): the call sequence of INTEGRATION.md section 2 / integration/hp3d_gpu_driver.F90 from C."""
import ctypes as C
import json
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("harness") / "harness")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-o", exe, os.path.join(HERE, "c", "harness.c"), "-ldl", "-lm"])
    return exe


def fortran_layout(src, tname):
    """(sizeof, {component: (offset, size)}) of a `type, bind(C) :: tname` block under the C interoperability rules."""
    m = re.search(r"type,\s*bind\(C\)\s*::\s*%s\b(.*?)end type" % tname, src, re.S | re.I)
    assert m, tname
    params = {k: int(v) for k, v in re.findall(r"(\w+)\s*=\s*(\d+)", " ".join(re.findall(r"integer\(c_int\),\s*parameter\s*::(.*)", src)))}
    off, align, out = 0, 1, {}
    for line in m.group(1).splitlines():
        line = line.split("!")[0].strip()
        if not line:
            continue
        kind, names = line.split("::")
        sz = {"integer(c_int)": 4, "real(c_double)": 8}[kind.strip()]
        for nm, dim in re.findall(r"(\w+)(?:\((\w+)\))?", names):
            n = 1 if not dim else (int(dim) if dim.isdigit() else params[dim])
            off = (off + sz - 1) // sz * sz
            out[nm] = (off, sz * n)
            off += sz * n
            align = max(align, sz)
    return (off + align - 1) // align * align, out


def test_struct_layouts_agree(harness):
    lay = json.loads(subprocess.check_output([harness, "layout"]))
    f90 = open(os.path.join(ROOT, "integration", "hp3d_gpu_mod.F90")).read()
    for tname, key in (("hp3d_params", "params"), ("hp3d_physics", "physics")):
        size, comps = fortran_layout(f90, tname)
        assert size == lay["sizeof_" + key], (tname, size, lay["sizeof_" + key])
        assert {k: tuple(v) for k, v in lay[key].items()} == comps, tname
    from hp3d_b200 import _lib, api
    assert C.sizeof(_lib.Params) == lay["sizeof_params"] and C.sizeof(api.Physics) == lay["sizeof_physics"]
    for nm, (o, s) in lay["params"].items():
        fld = getattr(_lib.Params, nm)
        assert (fld.offset, fld.size) == (o, s), nm
    for nm, (o, s) in lay["physics"].items():
        fld = getattr(api.Physics, nm)
        assert (fld.offset, fld.size) == (o, s), nm


def test_driver_binds_what_it_calls():
    """every hp3d_gpu_* routine integration/hp3d_gpu_driver.F90 calls has an interface in hp3d_gpu_mod.F90, with as many arguments"""
    mod = open(os.path.join(ROOT, "integration", "hp3d_gpu_mod.F90")).read()
    drv = open(os.path.join(ROOT, "integration", "hp3d_gpu_driver.F90")).read()
    drv = "\n".join(ln.split("!")[0] for ln in drv.splitlines() if not ln.lstrip().startswith("!"))
    drv = re.sub(r"&\s*\n\s*", " ", drv)
    mod = re.sub(r"&\s*\n\s*", " ", "\n".join(ln.split("!")[0] for ln in mod.splitlines()))
    iface = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()]) for m in re.finditer(r"(?:function|subroutine)\s+(hp3d_gpu_\w+)\s*\(([^)]*)\)", mod)}
    own = set(re.findall(r"subroutine\s+(hp3d_gpu_\w+)", drv))

    def nargs(s, start):   # arguments of the call whose '(' is at start
        depth, n, i, seen = 0, 0, start, False
        while True:
            ch = s[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    return n + (1 if seen else 0)
            elif ch == "," and depth == 1:
                n += 1
            elif depth >= 1 and not ch.isspace():
                seen = True
            i += 1

    used = 0
    for m in re.finditer(r"\b(hp3d_gpu_\w+)\s*\(", drv):
        name = m.group(1)
        if name in own or name == "hp3d_gpu_check":
            continue
        assert name in iface, name
        assert nargs(drv, m.end() - 1) == iface[name], (name, nargs(drv, m.end() - 1), iface[name])
        used += 1
    assert used >= 8


@pytest.mark.gpu
def test_replay_integration_sequence_from_c(harness, gpu):
    from hp3d_b200 import _lib
    out = subprocess.run([harness, "replay", _lib.LIB_PATH, "3", "5"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    r = json.loads(out.stdout)
    assert r["hermitian_defect"] == 0.0 and r["cloc_vs_host_aii"] == 0.0 and r["resident"] == 5 and r["spilled"] == 0
    assert r["bwd_err"] <= 1e-12 * (1 + r["bwd_max"])
