"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/g2048.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import g2048
from g2048 import _lib


def test_library_builds_and_exports_every_declared_symbol():
    path = g2048.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    declared = g2048.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, missing
    assert set(_lib._SIG) == set(declared), set(_lib._SIG) ^ set(declared)
    assert handle.g2048_version() == 200


def test_only_the_c_abi_is_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", g2048.build()], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert syms and all(s.startswith("g2048_") for s in syms), [s for s in syms if not s.startswith("g2048_")][:5]


def test_sass_is_sm100a_with_bulk_copy_and_256bit_loads():
    sass = subprocess.run(["cuobjdump", "-sass", g2048.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass          # cp.async.bulk (TMA) staging of the row LUT
    assert ".256" in sass            # one 32-byte load per Q-table slot
    assert "ATOMG.E.CAS" in sass     # key insertion and the atomic q <- q + lr (target - q)


@pytest.mark.skipif(ctypes.CDLL(g2048.build()).g2048_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback():
    with pytest.raises(g2048.G2048Error, match="no CPU fallback"):
        g2048.Game2048_env()
    with pytest.raises(g2048.G2048Error):
        g2048.init(0)
    L = g2048.lib()
    b = np.zeros(4, np.uint64)
    rc = L.g2048_legal_mask(b.ctypes.data, b.ctypes.data, 4, None)   # compute entry without init -> error code
    assert rc != 0 and L.g2048_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2048_q-learning_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "g2048_oracle" not in text, f


def test_tile_packing_roundtrip():
    rng = np.random.RandomState(0)
    for _ in range(200):
        lv = rng.randint(0, 16, size=16)
        tiles = np.where(lv > 0, 1 << lv.astype(np.int64), 0).reshape(4, 4)
        b = g2048.pack_tiles(tiles)
        assert np.array_equal(g2048.unpack_tiles(b), tiles)
    with pytest.raises(ValueError):
        g2048.pack_tiles(np.full((4, 4), 3))


def test_header_compiles_as_c_and_links():
    """include/g2048.h is plain C (C99) and a C program links against libg2048.so without any C++/CUDA symbols."""
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so_dir = os.path.dirname(g2048.build())
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "demo")
        res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                              os.path.join(root, "examples", "c_api_demo.c"), "-L" + so_dir, "-lg2048",
                              "-Wl,-rpath," + so_dir, "-o", exe], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        run = subprocess.run([exe], capture_output=True, text=True)
        if ctypes.CDLL(g2048.build()).g2048_device_count() == 0:
            assert run.returncode == 2 and "no CPU fallback" in run.stderr      # refuses loudly without a GPU


def test_ctypes_signatures_agree_with_the_header():
    """Every prototype of include/g2048.h against the ctypes table of _lib.py: same number of arguments, pointers bound
    as void pointers, 64-bit integers as 64-bit, floats/doubles as such, and the same kind of return value."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "g2048.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    protos = re.findall(r"G2048_API\s+([^;{]+?)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) == len(g2048.declared_symbols())

    def kind(decl):
        decl = " ".join(decl.split())
        if "*" in decl:
            return "ptr"
        base = decl.rsplit(" ", 1)[0] if " " in decl else decl
        base = base.replace("const ", "").strip()
        return {"int": "i32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "size_t": "u64",
                "float": "f32", "double": "f64", "void": "void"}[base]

    ctype_kind = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int: "i32", ctypes.c_int32: "i32",
                  ctypes.c_uint32: "u32", ctypes.c_int64: "i64", ctypes.c_uint64: "u64", ctypes.c_size_t: "u64",
                  ctypes.c_float: "f32", ctypes.c_double: "f64", None: "void"}
    for ret_and_name, params in protos:
        parts = ret_and_name.split()
        name = parts[-1].lstrip("*")
        ret = " ".join(parts[:-1]) + ("*" if parts[-1].startswith("*") or parts[-2].endswith("*") else "")
        restype, argtypes = _lib._SIG[name]
        plist = [p for p in (q.strip() for q in params.split(",")) if p and p != "void"]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes))
        for i, (p, a) in enumerate(zip(plist, argtypes)):
            assert kind(p + " x" if " " not in p else p) == ctype_kind[a], (name, i, p, a)
        assert kind(ret + " x") == ctype_kind[restype], (name, ret, restype)


def test_routed_buffer_size_is_a_pure_host_function():
    """g2048_routed_buffer_bytes needs no device: 0 for bad arguments, 256-byte granular, grows with world and cap, and
    holds what the header says (requests 2 x 8 B, answers 2 x (8 + 16) B per env and peer, records 8 B per env and peer)."""
    L = _lib.lib()
    f = L.g2048_routed_buffer_bytes
    assert f(0, 100) == 0 and f(17, 100) == 0 and f(2, 0) == 0
    a, b, c = f(2, 1 << 20), f(8, 1 << 20), f(8, 1 << 21)
    assert a % 256 == 0 and a < b < c
    per_env_and_peer = 2 * 8 + 2 * 8 + 2 * 16 + 8
    assert abs(b - 1024 - 8 * (1 << 20) * per_env_and_peer) < 8 * 5 * 256
