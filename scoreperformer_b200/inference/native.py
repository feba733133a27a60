"""Loader / builder of the host-side C library of the rendering path (include/spb200_host.h).

`libspb200_host.so` holds the onset recurrence of the SPMuple2 messenger (csrc/onset_times.c).  It is built in-tree with gcc
(`build()`, called by `__graft_entry__.build()`), git-ignored like the CUDA library and shipped with the snapshot.  The messenger
uses it when it is present and `SPB_HOST_NATIVE` is not "0"; the numpy statement of the same recurrence stays in messengers.py as
the definition the C code is tested against (tests/test_inference_host.py: identical bits on every messenger / rendering golden).
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import shutil
import subprocess
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "csrc", "onset_times.c")
HEADER = os.path.join(ROOT, "include", "spb200_host.h")
LIB_PATH = os.path.join(HERE, "csrc", "libspb200_host.so")
CFLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-shared", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "include")]

_handle: Optional[ctypes.CDLL] = None
_tried = False


def _digest() -> str:
    h = hashlib.sha256(" ".join(CFLAGS).encode())
    for path in (SRC, HEADER):
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    stamp = LIB_PATH + ".srchash"
    want = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB_PATH
    if shutil.which("gcc") is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH
        raise RuntimeError("gcc not found and no prebuilt libspb200_host.so")
    res = subprocess.run(["gcc"] + CFLAGS + ["-o", LIB_PATH, SRC, "-lm"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed on onset_times.c:\n" + res.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(want)
    global _handle, _tried
    _handle, _tried = None, False
    return LIB_PATH


def lib() -> Optional[ctypes.CDLL]:
    """The loaded library, or None when it has not been built / is switched off."""
    global _handle, _tried
    if os.environ.get("SPB_HOST_NATIVE", "1") == "0":
        return None
    if _tried:
        return _handle
    _tried = True
    if os.path.exists(LIB_PATH):
        h = ctypes.CDLL(LIB_PATH)
        h.spb_host_abi_version.restype = ctypes.c_int
        d, i64, u8, ci = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int
        h.spb_host_onset_times.restype = ctypes.c_int
        h.spb_host_onset_times.argtypes = [ci, d, d, d, d, d, u8, i64, ci, i64, d, ci, d, ci, ctypes.c_double, ctypes.c_double, ci, ci,
                                           ctypes.c_double, ctypes.c_double, ci, ci, d, ci, d, d, ctypes.POINTER(ci), ctypes.POINTER(ci),
                                           ctypes.POINTER(ci)]
        if h.spb_host_abi_version() >= 1:
            _handle = h
    return _handle
