"""Minimal reader for R's RDX2 (XDR) serialisation — enough for the reference's data/*.RData.

TEST INFRASTRUCTURE: used only by tests/golden/make_golden.py (run in the build container where
/root/reference exists) to decode the reference's bundled fixtures; nothing at run time reads
/root/reference.  Handles pairlists, lists, numeric/integer/logical/string vectors, attributes,
symbols with the reference table, and factors as plain ints.
"""
from __future__ import annotations

import bz2
import gzip
import struct

import numpy as np


class _Reader:
    def __init__(self, buf: bytes):
        self.b, self.p, self.refs = buf, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def take(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 254:  # NILVALUE
            return None
        if t == 255:  # REFSXP
            return self.refs[(flags >> 8) - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:  # LISTSXP (pairlist) -> dict / list
            out = {}
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out[tag if tag is not None else len(out)] = car
                flags = self.i32()
                t = flags & 0xFF
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
                if t == 254:
                    break
                assert t == 2, t
            return out
        if t == 9:  # CHARSXP
            n = self.i32()
            return None if n == -1 else self.take(n).decode("utf8", "replace")
        if t in (10, 13):  # LGLSXP, INTSXP
            n = self.i32()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif t == 14:  # REALSXP
            n = self.i32()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif t == 16:  # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif t == 19:  # VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError(f"SEXP type {t} at {self.p}")
        attrs = self.item() if has_attr else None
        return _wrap(v, attrs)


class RObj:
    """value + attributes (names, dim, class, row.names ...)"""

    def __init__(self, value, attrs):
        self.value, self.attrs = value, attrs or {}

    @property
    def names(self):
        n = self.attrs.get("names")
        return unwrap(n) if n is not None else None

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.value[self.names.index(key)]
        return self.value[key]

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attrs={list(self.attrs)})"


def _wrap(v, attrs):
    return RObj(v, attrs) if attrs else v


def unwrap(x):
    return x.value if isinstance(x, RObj) else x


def matrix(x) -> np.ndarray:
    """numeric RObj with a dim attribute -> (rows, cols) array"""
    dim = unwrap(x.attrs["dim"])
    return np.asarray(x.value).reshape((int(dim[0]), int(dim[1])), order="F")


def read_rdata(path: str) -> dict:
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    assert raw[:5] == b"RDX2\n", raw[:5]
    assert raw[5:7] == b"X\n"
    rd = _Reader(raw)
    rd.p = 7
    rd.i32(), rd.i32(), rd.i32()  # format version, writer R version, min reader version
    return rd.item()
