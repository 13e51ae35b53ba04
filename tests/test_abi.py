"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads without a GPU, exports every
symbol include/ipm_b200.h declares, the ctypes table agrees with the header, and the product path fails loudly
(no CPU fallback) when no B200 is visible."""

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ipm_b200.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(ipm_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from ipm_b200 import _abi

    return _abi.lib()


def test_header_symbols_are_exported(lib):
    fns = header_functions()
    assert len(fns) >= 30
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in include/ipm_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    from ipm_b200 import _abi

    fns = header_functions()
    assert set(_abi.SIGNATURES) == set(fns)
    for name, (_, argtypes) in _abi.SIGNATURES.items():
        assert len(argtypes) == fns[name], f"{name}: ctypes has {len(argtypes)} args, header {fns[name]}"


def test_library_is_sm100a_only():
    import subprocess

    from ipm_b200 import _abi

    out = subprocess.run(["cuobjdump", "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_blackwell_evidence_in_sass():
    """TMA (UTMALDG) and FP64 tensor (DMMA) instructions are what the hot kernels are made of."""
    import subprocess

    from ipm_b200 import _abi

    sass = subprocess.run(["cuobjdump", "-sass", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass and "DMMA" in sass and "SYNCS" in sass


def test_no_device_is_loud(lib):
    import torch

    from ipm_b200 import _abi

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    assert lib.ipm_device_ok() == _abi.IPM_ERR_NO_DEVICE
    with pytest.raises(_abi.IpmError):
        _abi.require_device()
    import numpy as np

    from ipm_b200.LPSolver import LPSolver

    with pytest.raises(_abi.IpmError):  # the drop-in class must not silently fall back to NumPy
        LPSolver(c=np.ones(3), C=np.eye(3), d=np.ones(3), check_cvxpy=False, suppress_print=True)


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "interiorpoint-gpu_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_every_call_site_passes_the_declared_number_of_arguments():
    """Static check of the host code: every `L("ipm_...", ...)` (Launcher call: the stream is appended) and
    `_abi.call("ipm_...", ...)` in the package passes as many positional arguments as the ctypes table declares.
    ctypes only complains at run time on a GPU box; this catches arity slips on the CPU."""
    import ast
    import glob
    import os

    from ipm_b200 import _abi

    pkg = os.path.dirname(_abi.__file__)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(pkg, "*.py")) + glob.glob(os.path.join(root, "tools", "*.py")) + glob.glob(
        os.path.join(root, "tests", "*.py")) + [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    checked = 0
    for path in sorted(files):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            if not isinstance(node, ast.Call) or not node.args:
                continue
            first = node.args[0]
            if not (isinstance(first, ast.Constant) and isinstance(first.value, str) and first.value.startswith("ipm_")):
                continue
            if any(isinstance(a, ast.Starred) for a in node.args):
                continue  # forwarded argument lists are checked where they are built
            name = first.value
            f = node.func
            is_abi_call = isinstance(f, ast.Attribute) and f.attr == "call"
            is_launcher = (isinstance(f, ast.Name) and f.id == "L") or (isinstance(f, ast.Attribute) and f.attr == "L")
            if not (is_abi_call or is_launcher):
                continue  # some other function that happens to take an "ipm_..." string
            assert name in _abi.SIGNATURES, f"{path}:{node.lineno}: {name} is not in the ctypes table"
            want = len(_abi.SIGNATURES[name][1]) - (0 if is_abi_call else 1)
            got = len(node.args) - 1
            assert got == want, f"{path}:{node.lineno}: {name} takes {want} arguments here, {got} given"
            checked += 1
    assert checked > 40


def test_definitions_match_the_header():
    """The .cu files do not include the public header (they are compiled as C++ with their own helpers), so nothing
    but this test ties the parameter list of every `extern "C"` definition to its prototype in include/ipm_b200.h:
    same number of parameters, same parameter types in the same order."""
    import glob
    import os

    def norm(params):
        out = []
        for p in params.split(","):
            p = re.sub(r"\b(const|__restrict__)\b", " ", p)
            p = re.sub(r"\s+", " ", p).strip()
            m = re.match(r"^(.*?)(\w+)$", p)  # drop the parameter name
            out.append(re.sub(r"\s+", "", m.group(1)) if m and m.group(1).strip() else re.sub(r"\s+", "", p))
        return out

    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {m.group(1): norm(m.group(2)) for m in re.finditer(r"\b(ipm_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
              if m.group(2).strip() not in ("", "void")}
    csrc = os.path.join(os.path.dirname(HEADER), "..", "interiorpoint-gpu_b200", "csrc")
    seen = set()
    for path in glob.glob(os.path.join(csrc, "*.cu")):
        src = re.sub(r"//[^\n]*", "", open(path).read())
        for m in re.finditer(r'extern "C"\s+[\w\s\*]*?\b(ipm_\w+)\s*\(([^{;]*?)\)\s*\{', src, flags=re.S):
            name, params = m.group(1), m.group(2)
            if name not in protos:
                continue  # library-internal helpers are not in the header
            assert norm(params) == protos[name], f"{os.path.basename(path)}: {name}: {norm(params)} vs header {protos[name]}"
            seen.add(name)
    assert seen == set(protos), sorted(set(protos) - seen)


def test_ctypes_argument_types_match_the_header():
    """Same table, types this time: a c_int where the prototype says long long (or a double passed as an int) would be
    silently misread on the device side."""
    import ctypes as C

    from ipm_b200 import _abi

    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    scalars = {"int": C.c_int, "double": C.c_double, "longlong": C.c_longlong, "unsignedlonglong": C.c_ulonglong,
               "unsignedint": C.c_uint, "unsigned": C.c_uint}
    for m in re.finditer(r"\b(ipm_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        if params in ("", "void"):
            continue
        argtypes = _abi.SIGNATURES[name][1]
        for k, p in enumerate(params.split(",")):
            p = re.sub(r"\bconst\b", " ", p)
            mm = re.match(r"^(.*?)(\w+)\s*$", p.strip())
            ctype = re.sub(r"\s+", "", mm.group(1))
            got = argtypes[k]
            if ctype.endswith("**"):
                want = {C.POINTER(C.c_void_p)}
            elif ctype.endswith("*"):
                # scalar outputs on the HOST are bound as typed pointers, device pointers as void*
                want = {C.c_void_p, C.POINTER(scalars.get(ctype[:-1], C.c_void_p))}
            else:
                want = {scalars[ctype]}
            assert got in want, f"{name} argument {k} ({p.strip()}): ctypes {got}, header {ctype}"
