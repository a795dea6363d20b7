#!/usr/bin/env python
"""Decode the reference's bundled fixture into a small .npz the tests can carry.

Source : /root/reference/data/EXAMPLE_DATA_list.rda  (doc: R/EXAMPLE_DATA_list.r:1-18)
Format : bzip2 -> "RDX2\\nX\\n" (R serialisation, XDR, version 2) -> pairlist
         { EXAMPLE_DATA_list = VECSXP(4) of REALSXP with dim/dimnames/names }.
Output : tests/golden/example_data.npz  (inputdata 20x74 float64, inputbeta 74,
         mu 20, size 20, gene/cell names).

Only runs where /root/reference exists (the build container); the .npz it
writes is committed, so nothing on the GPU box reads the reference tree.
No R is needed: this is a ~80-line reader for the handful of SEXP types the
file contains (LISTSXP 2, SYMSXP 1, CHARSXP 9, INTSXP 13, REALSXP 14, STRSXP 16,
VECSXP 19, NILVALUE 254, REFSXP 255).
"""
import bz2
import struct
import sys
from pathlib import Path

import numpy as np


class _XdrReader:
    def __init__(self, buf):
        self.b = buf
        self.o = 0
        self.syms = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def f64n(self, n):
        a = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.o).astype(np.float64)
        self.o += 8 * n
        return a

    def i32n(self, n):
        a = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.o).astype(np.int32)
        self.o += 4 * n
        return a

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == 254:          # NILVALUE_SXP
            return None
        if typ == 255:          # REFSXP
            return self.syms[(flags >> 8) - 1]
        if typ == 1:            # SYMSXP
            name = self.item()
            self.syms.append(name)
            return name
        if typ == 9:            # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.o:self.o + n].decode("latin-1")
            self.o += n
            return s
        if typ == 2:            # LISTSXP (pairlist) -> dict tag -> value
            out = {}
            while True:
                attr = self.item() if has_attr else None  # noqa: F841
                tag = self.item() if has_tag else None
                car = self.item()
                out[tag] = car
                flags = self.i32()
                typ = flags & 0xFF
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
                if typ == 254:
                    return out
                if typ != 2:
                    raise ValueError("unexpected cdr type %d" % typ)
        if typ in (13, 14, 16, 19):
            n = self.i32()
            if typ == 13:
                val = self.i32n(n)
            elif typ == 14:
                val = self.f64n(n)
            elif typ == 16:
                val = [self.item() for _ in range(n)]
            else:
                val = [self.item() for _ in range(n)]
            attrs = self.item() if has_attr else None
            return {"value": val, "attr": attrs} if attrs else {"value": val, "attr": {}}
        raise ValueError("unsupported SEXP type %d at offset %d" % (typ, self.o))


def decode(path):
    raw = bz2.decompress(Path(path).read_bytes())
    assert raw[:5] == b"RDX2\n", raw[:8]
    assert raw[5:7] == b"X\n"
    r = _XdrReader(raw)
    r.o = 7
    version, writer, minver = r.i32(), r.i32(), r.i32()
    assert version == 2, version
    top = r.item()
    lst = top["EXAMPLE_DATA_list"]
    names = [s for s in lst["attr"]["names"]["value"]]
    out = {}
    for nm, el in zip(names, lst["value"]):
        v = el["value"]
        a = el["attr"]
        if "dim" in a:
            d = a["dim"]["value"]
            v = v.reshape((int(d[1]), int(d[0]))).T.copy()   # column-major -> (rows, cols)
            dn = a["dimnames"]["value"]
            out[nm + "_rownames"] = np.array(dn[0]["value"])
            out[nm + "_colnames"] = np.array(dn[1]["value"])
        elif "names" in a:
            out[nm + "_names"] = np.array(a["names"]["value"])
        out[nm] = v
    return out


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/EXAMPLE_DATA_list.rda"
    d = decode(src)
    dst = Path(__file__).with_name("example_data.npz")
    np.savez_compressed(dst, **d)
    for k, v in d.items():
        print(k, getattr(v, "shape", None), getattr(v, "dtype", None))
    x = d["inputdata"]
    print("inputdata max %g mean %.4f zeros %.3f" % (x.max(), x.mean(), (x == 0).mean()))
    print("inputbeta min %.4f max %.4f mean %.4f" % (d["inputbeta"].min(), d["inputbeta"].max(), d["inputbeta"].mean()))
    print("mu", d["mu"].min(), d["mu"].max(), "size", d["size"].min(), d["size"].max())
