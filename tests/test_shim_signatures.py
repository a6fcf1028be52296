"""The Julia shim cannot run in the build container (no Julia), so its `ccall` / `@threadcall` signatures are checked
statically against the prototypes of include/psa_gpu.h: same entry point names, same number of arguments, and every
Julia argument type compatible with the C parameter type.  The same check runs over the binding text of
INTEGRATION.md."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Julia ccall type -> the C parameter types it may be bound to
COMPAT = {
    "Ptr{Cvoid}": {"void*"},
    "Ref{Ptr{Cvoid}}": {"void**"},
    "Cint": {"int"},
    "Int64": {"int64_t"},
    "Float64": {"double"},
    "Cstring": {"const char*"},
    "Ptr{Float64}": {"const double*", "double*"},
    "Ptr{ComplexF64}": {"const double*", "double*"},  # complex128 arrays travel as interleaved doubles
    "Ref{Float64}": {"double*"},
    "Ptr{Int32}": {"const int32_t*", "int32_t*", "const int*", "int*"},
    "Ptr{Int64}": {"int64_t*"},
    "Ptr{Cint}": {"int*", "const int*", "volatile int*"},
    "Ref{Cint}": {"int*"},
}
RET = {"Cint": "int", "Cstring": "const char*"}


def _c_prototypes():
    hdr = open(os.path.join(ROOT, "include", "psa_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"^\s*#.*$", "", hdr, flags=re.M)  # preprocessor lines
    protos = {}
    for ret, name, params in re.findall(r"^((?:const\s+)?\w+\s*\**)\s*(psa_gpu_\w+)\s*\(([^)]*)\)\s*;", hdr, flags=re.M):
        types = []
        for p in [q.strip() for q in params.split(",")]:
            if p in ("void", ""):
                continue
            p = re.sub(r"\s+", " ", p)
            m = re.match(r"(.*?)(\w+)$", p)          # strip the parameter name
            t = m.group(1).strip() if m and m.group(1).strip() else p
            types.append(t.replace(" *", "*").replace("* ", "*"))
        protos[name] = (re.sub(r"\s+", " ", ret).strip().replace(" *", "*"), types)
    return protos


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{([":
            depth += 1
        elif ch in "})]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _julia_calls(text):
    """(name, return type, [argument types], number of actual arguments) of every ccall / @threadcall in `text`."""
    calls = []
    for m in re.finditer(r"(?:ccall|@threadcall)\(\(:(psa_gpu_\w+), libpsagpu\),\s*(\w+),\s*\(", text):
        i, depth = m.end(), 1
        while depth:  # the tuple of argument types
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        types = _split_top(text[m.end():i - 1])
        j, depth = i, 1  # the rest of the call: the actual arguments
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        args = _split_top(text[i:j - 1].lstrip(", \n"))
        calls.append((m.group(1), m.group(2), types, len(args)))
    return calls


@pytest.mark.parametrize("path", ["julia/PseudospectraGPU.jl", "INTEGRATION.md"])
def test_julia_bindings_match_the_header(path):
    protos = _c_prototypes()
    calls = _julia_calls(open(os.path.join(ROOT, path)).read())
    assert len(calls) >= 6, "no bindings found in " + path
    for name, ret, types, nargs in calls:
        assert name in protos, name
        cret, ctypes_ = protos[name]
        assert RET[ret] == cret, (name, ret, cret)
        assert len(types) == len(ctypes_) == nargs, (name, len(types), len(ctypes_), nargs)
        for k, (jt, ct) in enumerate(zip(types, ctypes_)):
            assert jt in COMPAT and ct in COMPAT[jt], (name, k, jt, ct)


def test_shim_binds_the_whole_compute_boundary():
    """Every entry point a Julia caller needs for psa_compute / zoom / modeplot is bound by the shim."""
    names = {c[0] for c in _julia_calls(open(os.path.join(ROOT, "julia", "PseudospectraGPU.jl")).read())}
    assert {"psa_gpu_init", "psa_gpu_free", "psa_gpu_set_thresholds", "psa_gpu_set_matrix_ex", "psa_gpu_grid_batched",
            "psa_gpu_postprocess", "psa_gpu_project", "psa_gpu_pseudomode", "psa_gpu_strerror"} <= names
