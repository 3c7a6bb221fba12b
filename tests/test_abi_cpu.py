"""CPU-side checks of the drop-in boundary (no GPU needed):
  * libsfhcuda.so loads and exports every symbol include/sfhcuda.h declares;
  * the ctypes table in the host package covers exactly those symbols;
  * without a device every compute entry point fails loudly with SFH_ERR_NO_DEVICE (no CPU fallback);
  * the product package never imports / links / loads anything under oracle/.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "starformationhistories.jl_b200")
HEADER = os.path.join(ROOT, "include", "sfhcuda.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfh_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import sfh_b200
    syms = header_symbols()
    assert len(syms) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", sfh_b200._lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (sfh_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared in sfhcuda.h but not exported: {missing}"
    assert set(sfh_b200._lib.PROTOTYPES) == set(syms), set(sfh_b200._lib.PROTOTYPES) ^ set(syms)
    for s in syms:
        getattr(sfh_b200._lib.lib, s)


def test_struct_layouts_match_header():
    import sfh_b200
    L = sfh_b200._lib
    assert C.sizeof(L.sfh_opts) == 56      # 2*i32, 2*i64, f64, 6*i32
    assert C.sizeof(L.sfh_stats) == 24
    assert L.lib.sfh_version() == 1


def test_every_struct_field_offset_matches_header(tmp_path):
    """Every ctypes mirror in _lib.py has the size and the field offsets gcc computes from include/sfhcuda.h itself
    (a field added to the header and not to the mirror -- or the other way round -- fails here, on the CPU)."""
    import sfh_b200
    L = sfh_b200._lib
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    names = re.findall(r"typedef\s+struct\s+(sfh_\w+)\s*\{", src)
    assert len(names) >= 9
    mirrors = {n: getattr(L, n) for n in names}      # AttributeError = a struct of the header without a mirror
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sfhcuda.h"', 'int main(void) {']
    for n, cls in mirrors.items():
        lines.append(f'  printf("{n} . %zu\\n", sizeof({n}));')
        for f in cls._fields_:
            lines.append(f'  printf("{n} {f[0]} %zu\\n", offsetof({n}, {f[0]}));')
    lines += ['  return 0;', '}']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.dirname(HEADER), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        n, f, v = line.split()
        want = C.sizeof(mirrors[n]) if f == "." else getattr(mirrors[n], f).offset
        assert int(v) == want, f"{n}.{f}: header {v}, ctypes {want}"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in mirrors.values())
    # and the header holds no field the mirror lacks: count the declarators of each struct body
    for n, cls in mirrors.items():
        body = re.search(r"typedef\s+struct\s+" + n + r"\s*\{(.*?)\}\s*" + n + r"\s*;", src, flags=re.S).group(1)
        decls = sum(len(stmt.split(",")) for stmt in body.split(";") if stmt.strip())
        assert decls == len(cls._fields_), f"{n}: header declares {decls} fields, ctypes mirrors {len(cls._fields_)}"


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sfh_b200
    assert sfh_b200.device_count() == 0
    with pytest.raises(sfh_b200.SFHError) as ei:
        sfh_b200.DeviceStack(np.ones((8, 3)), np.ones(8))
    assert ei.value.status == sfh_b200._lib.SFH_ERR_NO_DEVICE
    with pytest.raises(sfh_b200.SFHError):
        sfh_b200.fg_(True, np.zeros(3), np.ones(3), np.ones((8, 3)), np.ones(8))
    with pytest.raises(ValueError):       # shape errors are raised before anything else
        sfh_b200.DeviceStack(np.ones((8, 3)), np.ones(7))


def test_argument_validation_without_device():
    import sfh_b200
    L = sfh_b200._lib
    assert L.lib.sfh_device_count(None) == L.SFH_ERR_INVALID_ARG
    assert b"NULL" in L.lib.sfh_last_error()
    h = C.c_void_p()
    assert L.lib.sfh_stack_create(C.byref(h), None, 4, 2, L.SFH_F64, None, L.SFH_F64, None) == L.SFH_ERR_INVALID_ARG
    assert L.lib.sfh_stack_destroy(None) == L.SFH_OK       # idempotent on NULL (finalizer-safe)
    assert L.lib.sfh_ctx_destroy(None) == L.SFH_OK


def test_product_never_touches_oracle():
    bad = []
    for dp, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"\boracle\b|sfho_|libsfhoracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, f"product sources reference the oracle: {bad}"
    import sfh_b200
    out = subprocess.run(["ldd", sfh_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    hdr = open(HEADER).read()
    assert "oracle" not in hdr.lower()


def test_host_model_api_matches_reference_doctests():
    """mzr.jl:233-261, amr.jl:213-247, dispersion_models.jl:51-78 doctests on the host model classes."""
    import sfh_b200 as S
    m = S.PowerLawMZR(1.0, -1, 6)
    assert m.nparams() == 2 and abs(m(1e7)) < 1e-15
    g = m.gradient(1e8)
    assert g[0] == pytest.approx(2.0) and g[1] == 1.0 and g[2] == pytest.approx(1 / 1e8 / np.log(10))
    assert S.PowerLawMZR(1.0, -1, 7, (True, False)).update_params((2.0, -2)) == S.PowerLawMZR(2.0, -2, 7, (True, False))
    assert m.transforms() == (1, 0) and S.PowerLawMZR(1.0, -1, 7, (True, False)).free_params() == (True, False)
    with pytest.raises(ValueError):
        S.PowerLawMZR(-1.0, -1)
    d = S.GaussianDispersion(0.2)
    assert d(1.0, 1.2) == pytest.approx(np.exp(-0.5))
    assert d.gradient(1.0, 1.2) == (pytest.approx(3.0326532985631656), pytest.approx(-3.0326532985631656))
    assert d.transforms() == (1,) and S.GaussianDispersion(0.2, (False,)).free_params() == (False,)
    with pytest.raises(ValueError):
        S.GaussianDispersion(0.0)
    assert S.LinearAMR(0.05, -1.6, 12).transforms() == (1, 0) and S.LogarithmicAMR(1e-4, 5e-5, 12).transforms() == (1, 1)
    assert S.nparams(m, d) == 3
    with pytest.raises(ValueError):
        S.LogarithmicAMR(1e-4, -1.0)


def test_host_calculate_coeffs_matches_oracle():
    import oracle as O
    import sfh_b200 as S
    from conftest import make_hier_problem
    p = make_hier_problem(nj=21, nk=26, nb=4, shuffle=True, ragged=True)
    for model, kind, fixed in [(S.PowerLawMZR(1.0, -2.0, 6.0), O.POWERLAW_MZR, (6.0,)),
                               (S.LinearAMR(0.05, -1.6, 12.0), O.LINEAR_AMR, (12.0,)),
                               (S.LogarithmicAMR(1e-4, 5e-5, 12.0), O.LOG_AMR, (12.0,))]:
        got = S.calculate_coeffs(model, S.GaussianDispersion(0.2), p["R"], p["logAge"], p["MH"])
        want = O.calculate_coeffs(kind, model.alpha, model.beta, fixed, 0.2, p["R"], p["logAge"], p["MH"])
        assert np.allclose(got, want, rtol=1e-13, atol=0)
    with pytest.raises(ValueError):
        S.calculate_coeffs(S.PowerLawMZR(1.0, -2.0), S.GaussianDispersion(0.2), p["R"][:-1], p["logAge"], p["MH"])


def test_identity_cache_detects_recycled_ids():
    """id() of a dead array can be reused by a new one: the stack cache must compare identities through weakrefs."""
    from sfh_b200 import fitting as F
    a = np.ones((4, 2)); d = np.ones(4)
    refs = F._weak_ids(a)
    assert F._same(refs, a) and not F._same(refs, np.ones((4, 2)))
    lst = [np.ones((2, 2)), np.zeros((2, 2))]
    r2 = F._weak_ids(lst)
    assert F._same(r2, lst) and not F._same(r2, [lst[0], np.zeros((2, 2))])
    del a
    assert refs[0]() is None and not F._same(refs, np.ones((4, 2)))
    assert F._weak_ids([[1, 2], [3, 4]]) is None


def test_plain_c_consumer(tmp_path):
    """include/sfhcuda.h is valid C99 and a plain C program links and runs against libsfhcuda.so (tests/abi_c_smoke.c)."""
    import sfh_b200
    exe = tmp_path / "abi_c_smoke"
    libdir = os.path.dirname(sfh_b200._lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_c_smoke.c"), "-o", str(exe), "-L", libdir, "-l:libsfhcuda.so", "-lm",
                    f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe), str(tmp_path / "smoke.sfh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "abi_c_smoke: ok" in r.stdout


def test_native_cpp_example_builds_and_runs(tmp_path):
    """examples/fit_sfh_native.cpp -- fit_sfh + tsample_sfh + persistence through nothing but the C-ABI -- compiles warning-free
    and, on a box without a GPU, says so and exits 0 (on a GPU box it runs a small problem end to end)."""
    import sfh_b200
    exe = tmp_path / "fit_sfh_native"
    libdir = os.path.dirname(sfh_b200._lib.LIB_PATH)
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "fit_sfh_native.cpp"), "-o", str(exe), "-L", libdir, "-l:libsfhcuda.so",
                    f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe), "40", "50", "8", "6", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    if sfh_b200.device_count() == 0:
        assert "no CUDA device" in r.stdout
    else:
        assert "wrote" in r.stdout and os.path.exists(tmp_path / "fit.sfh")


def test_julia_struct_mirrors_match_header():
    """julia/SFHCuda.jl cannot be executed here (no julia binary): its isbits mirrors of the option / report structs are compared
    statically, field by field (name, width, order), with the ctypes mirrors the test above checks against gcc."""
    import sfh_b200
    L = sfh_b200._lib
    jl = open(os.path.join(PKG, "julia", "SFHCuda.jl")).read()
    pairs = {"SfhOpts": L.sfh_opts, "BfgsOpts": L.sfh_bfgs_opts, "BfgsReport": L.sfh_bfgs_report, "NutsOpts": L.sfh_nuts_opts,
             "LbfgsbOpts": L.sfh_lbfgsb_opts, "LbfgsbReport": L.sfh_lbfgsb_report}
    jl_types = {"Int32": C.c_int32, "Int64": C.c_int64, "UInt64": C.c_uint64, "Float64": C.c_double}
    for jname, cls in pairs.items():
        m = re.search(r"^(?:mutable\s+)?struct\s+" + jname + r"\b(.*?)\bend\s*$", jl, flags=re.S | re.M)
        assert m, f"{jname} not found in SFHCuda.jl"
        fields = re.findall(r"(\w+)::(\w+)", m.group(1))
        assert [(n, jl_types[t]) for n, t in fields] == [(f[0], f[1]) for f in cls._fields_], jname


def _strip_julia(s):
    """Julia source without comments and string contents (enough for structural checks)."""
    out, i, n = [], 0, len(s)
    while i < n:
        if s.startswith("#=", i):
            i = s.index("=#", i) + 2
        elif s[i] == "#":
            j = s.find("\n", i)
            i = n if j < 0 else j
        elif s.startswith('"""', i):
            i = s.index('"""', i + 3) + 3
            out.append('""')
        elif s[i] == '"':
            j = i + 1
            while s[j] != '"':
                j += 2 if s[j] == "\\" else 1
            i = j + 1
            out.append('""')
        else:
            out.append(s[i])
            i += 1
    return "".join(out)


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        depth += ch in "({["
        depth -= ch in ")}]"
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def test_julia_binding_structure_and_ccalls_match_header():
    """julia/SFHCuda.jl is the drop-in binding itself and cannot run here: every `ccall` must name a symbol the header declares,
    with as many argument types as the C prototype has parameters; brackets and `function/struct/module/do ... end` blocks balance."""
    code = _strip_julia(open(os.path.join(PKG, "julia", "SFHCuda.jl")).read())
    for o, c in ("()", "[]", "{}"):
        assert code.count(o) == code.count(c), (o, code.count(o), code.count(c))
    openers = re.findall(r"(?<![\w.:@])(function|struct|if|while|let|do|begin|module|try|quote|macro)(?![\w!])", code)
    block_for = re.findall(r"^\s*for\b", code, flags=re.M)          # a `for` that starts a line is a loop; the others are generators
    ends = re.findall(r"(?<![\w.:@\[])end(?![\w!])", code)
    assert len(openers) + len(block_for) == len(ends), (len(openers), len(block_for), len(ends))
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)

    def c_class(param):            # pointer / 32-bit / 64-bit / double: what the calling convention distinguishes
        param = param.strip()
        if "*" in param or "(" in param:
            return "ptr"
        t = re.sub(r"\b\w+$", "", param).replace("const", "").strip() if re.search(r"\s", param) else param
        return {"int": "i32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "size_t": "u64",
                "double": "f64"}.get(t, "ptr" if t.endswith("_fn") else ("i32" if t.startswith("sfh_") else "?" + t))      # sfh_* by value: enums

    def j_class(t):
        t = t.strip()
        if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
            return "ptr"
        return {"Cint": "i32", "Int32": "i32", "UInt32": "u32", "Int64": "i64", "UInt64": "u64", "Csize_t": "u64",
                "Float64": "f64", "Cdouble": "f64"}.get(t, "?" + t)

    nparams, classes = {}, {}
    for m in re.finditer(r"\b(sfh_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        plist = [] if args in ("void", "") else _split_top(args)
        nparams[m.group(1)] = len(plist)
        classes[m.group(1)] = [c_class(a) for a in plist]
    seen = 0
    for m in re.finditer(r"ccall\(\(:(\w+),\s*libsfh\),\s*[\w{}]+,\s*\(", code):
        name, i, depth = m.group(1), m.end(), 1
        j = i
        while depth:
            depth += code[j] == "("
            depth -= code[j] == ")"
            j += 1
        assert name in nparams, f"ccall of {name}: not declared in sfhcuda.h"
        assert len(_split_top(code[i:j - 1])) == nparams[name], f"ccall of {name}: {len(_split_top(code[i:j - 1]))} argument types, C prototype has {nparams[name]}"
        jt = [j_class(t) for t in _split_top(code[i:j - 1])]
        assert jt == classes[name], f"ccall of {name}: argument classes {jt}, C prototype {classes[name]}"
        seen += 1
    assert seen >= 20 and seen == len(re.findall(r"ccall\(", code))


def test_ctypes_prototypes_match_header_parameter_classes():
    """Every ctypes prototype in _lib.py has as many parameters as the C declaration and the same class per parameter
    (pointer / 32-bit / 64-bit integer / double) -- a c_int where the header says int64_t would truncate silently."""
    import sfh_b200
    L = sfh_b200._lib
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)

    def c_class(param):
        param = param.strip()
        if "*" in param or "(" in param:
            return "ptr"
        t = re.sub(r"\b\w+$", "", param).replace("const", "").strip() if re.search(r"\s", param) else param
        return {"int": "i32", "int32_t": "i32", "uint32_t": "i32", "int64_t": "i64", "uint64_t": "i64", "size_t": "i64",
                "double": "f64"}.get(t, "ptr" if t.endswith("_fn") else ("i32" if t.startswith("sfh_") else "?" + t))

    def py_class(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or hasattr(t, "_flags_") and issubclass(t, C._CFuncPtr):
            return "ptr"
        if issubclass(t, C._Pointer):
            return "ptr"
        return {C.c_int: "i32", C.c_int32: "i32", C.c_uint32: "i32", C.c_int64: "i64", C.c_uint64: "i64", C.c_size_t: "i64",
                C.c_double: "f64"}.get(t, "?" + getattr(t, "__name__", str(t)))

    checked = 0
    for m in re.finditer(r"\b([\w\s\*]+?)\b(sfh_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        name, args = m.group(2), m.group(3).strip()
        want = [] if args in ("void", "") else [c_class(a) for a in _split_top(args)]
        res, argtypes = L.PROTOTYPES[name]
        got = [py_class(t) for t in argtypes]
        assert got == want, f"{name}: ctypes {got}, header {want}"
        ret = m.group(1).strip()
        assert (res is C.c_char_p) == ("char" in ret) and (res is L._int or "char" in ret or res is C.c_int), (name, ret, res)
        checked += 1
    assert checked == len(L.PROTOTYPES)


_NULL_SWEEP = r"""
import ctypes as C, sys, faulthandler
faulthandler.enable()
sys.path.insert(0, sys.argv[1])
import sfh_b200
L = sfh_b200._lib
def zero(t):
    if t in (C.c_void_p, C.c_char_p):
        return None
    if issubclass(t, (C._Pointer, C._CFuncPtr)):
        return C.cast(None, t)
    return t(0)
for name in sorted(L.PROTOTYPES):
    res, argtypes = L.PROTOTYPES[name]
    r = getattr(L.lib, name)(*[zero(t) for t in argtypes])
    print(name, r if res is not C.c_char_p else "str", flush=True)
print("SWEEP DONE")
"""


def test_every_entry_point_survives_null_arguments():
    """The boundary's error contract ("every call returns int status ... never throws"): each of the library's
    entry points, called with NULL pointers and zero sizes, hands back a status -- SFH_ERR_INVALID_ARG, or SFH_OK for the
    finalizer-safe destroy / close calls -- instead of crashing.  Run in a child process so that a crash is a test failure."""
    import sys
    r = subprocess.run([sys.executable, "-c", _NULL_SWEEP, ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SWEEP DONE" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    got = dict(line.split() for line in r.stdout.splitlines() if line and not line.startswith("SWEEP"))
    syms = header_symbols()
    assert sorted(got) == syms
    ok_on_null = {"sfh_stack_destroy", "sfh_ctx_destroy", "sfh_file_close", "sfh_group_destroy"}
    for name, val in got.items():
        if name == "sfh_last_error":
            continue
        if name == "sfh_version":
            assert val == "1"
        elif name in ok_on_null:
            assert val == "0", (name, val)
        else:
            assert val == "1", f"{name} returned {val} for NULL arguments (expected SFH_ERR_INVALID_ARG = 1)"
