"""ctypes binding of libdcvit.so (C ABI in include/dcvit.h).

There is deliberately no fallback: if the library is missing or a call fails, an
exception is raised.  The library is built in-tree by diverse_channel_vit_b200.build.
"""
from __future__ import annotations

import ctypes
import re
from ctypes import c_int, c_longlong, c_void_p, c_float, c_char_p
from pathlib import Path

import os

PKG_DIR = Path(__file__).resolve().parent
# DCV_LIB=<variant> loads a validation build (libdcvit_<variant>.so, see build.py VARIANTS); default: the shipped library
_VARIANT = os.environ.get("DCV_LIB", "")
LIB_PATH = PKG_DIR / (f"libdcvit_{_VARIANT}.so" if _VARIANT else "libdcvit.so")
HEADER_PATH = PKG_DIR.parent / "include" / "dcvit.h"


class DcvError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every function name declared in include/dcvit.h."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcv_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DcvError(
                f"{LIB_PATH} not found: build it with `python -m diverse_channel_vit_b200.build` "
                "(there is no CPU/PyTorch fallback for the DiChaViT hot path)"
            )
        _lib = ctypes.CDLL(str(LIB_PATH))
        _lib.dcv_last_error.restype = c_char_p
        _lib.dcv_launch_count.restype = c_longlong
        for name in declared_symbols():
            if not hasattr(_lib, name) and not _VARIANT:  # validation builds may predate / omit debug entry points
                raise DcvError(f"libdcvit.so does not export {name} (stale build?)")
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dcv_last_error().decode(errors="replace")
        raise DcvError(f"{what or 'dcv call'} failed (code {rc}): {msg}")


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(stream=None) -> c_void_p:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)


def launch_count() -> int:
    return int(lib().dcv_launch_count())
