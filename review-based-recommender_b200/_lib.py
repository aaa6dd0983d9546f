"""ctypes binding of csrc/librbr_b200.so — the C-ABI declared in include/rbr_b200.h.

The prototypes are parsed from the header itself, so the binding cannot drift from the declared ABI.
There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(ROOT, "include", "rbr_b200.h")
LIB_PATH = os.path.join(_HERE, "csrc", "librbr_b200.so")

_SCALARS = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


def _ctype(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_char_p if decl.replace(" ", "") == "constchar*" else ctypes.c_void_p
    base = decl.replace("const", "").strip()
    return _SCALARS[base]


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object], List[str]]]:
    """{symbol: (restype, [argtypes], [argnames])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(rbr_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                argtypes.append(_ctype(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_ctype(ret), argtypes, argnames)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()

    def load(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"rbr_b200: native library not built ({LIB_PATH}). Run `python -c 'import __graft_entry__ as g; "
                    f"g.build()'` (needs nvcc; sm_100a). There is no CPU or PyTorch fallback for the hot path.")
            dll = ctypes.CDLL(LIB_PATH)
            for name, (ret, argtypes, _) in self.protos.items():
                fn = getattr(dll, name)          # AttributeError here = header/library mismatch: fail loudly
                fn.restype = ret
                fn.argtypes = argtypes
            self._dll = dll
        return self._dll

    def __getattr__(self, name):
        return getattr(self.load(), name)

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.load().rbr_last_error()
            raise RuntimeError(f"rbr_b200: {what} failed (code {rc}): {msg.decode() if msg else ''}")


lib = _Lib()
