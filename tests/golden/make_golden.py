#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's only binary fixture.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py

Source: /root/reference/test/problem_data/random_polish_qp.jld2 -- the problem
and Mosek solution that test/polishing.jl:69-93 checks at atol=1e-3.  h5py is
not installed, so this is a minimal HDF5 (superblock v2, object header v2,
compact layouts only) walker; JLD2 stores SparseMatrixCSC as a compound
{m, n, colptr-ref, rowval-ref, nzval-ref} whose refs are object addresses.
"""
import os
import struct
import sys

import numpy as np

SRC = "/root/reference/test/problem_data/random_polish_qp.jld2"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "random_polish_qp.npz")


class H5:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        sb = self.b.find(b"\x89HDF\r\n\x1a\n")
        assert sb >= 0 and self.b[sb + 8] == 2, "need superblock v2"
        assert self.b[sb + 9] == 8 and self.b[sb + 10] == 8
        self.base, _ext, _eof, self.root = struct.unpack_from("<QQQQ", self.b, sb + 12)

    def messages(self, addr):
        """Yield (type, payload bytes) of the v2 object header at relative address addr."""
        o = self.base + addr
        b = self.b
        assert b[o:o + 4] == b"OHDR" and b[o + 4] == 2, (o, b[o:o + 8])
        flags = b[o + 5]
        o += 6
        if flags & 0x20:
            o += 16
        if flags & 0x10:
            o += 4
        szlen = 1 << (flags & 3)
        size = int.from_bytes(b[o:o + szlen], "little")
        o += szlen
        chunks = [(o, size)]
        while chunks:
            o, size = chunks.pop(0)
            end = o + size
            while o + 4 <= end:
                mtype = b[o]
                msize = struct.unpack_from("<H", b, o + 1)[0]
                o += 4
                if flags & 0x04:
                    o += 2
                payload = b[o:o + msize]
                o += msize
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", payload, 0)
                    co = self.base + caddr
                    assert b[co:co + 4] == b"OCHK"
                    chunks.append((co + 4, clen - 8))
                elif mtype != 0:
                    yield mtype, payload

    def links(self, addr):
        out = {}
        for t, p in self.messages(addr):
            if t != 6:
                continue
            ver, fl = p[0], p[1]
            o = 2
            ltype = 0
            if fl & 0x08:
                ltype = p[o]; o += 1
            if fl & 0x04:
                o += 8
            if fl & 0x10:
                o += 1
            nlen_sz = 1 << (fl & 3)
            nlen = int.from_bytes(p[o:o + nlen_sz], "little"); o += nlen_sz
            name = p[o:o + nlen].decode(); o += nlen
            if ltype == 0:
                out[name] = struct.unpack_from("<Q", p, o)[0]
        return out

    def dataset(self, addr):
        """Return (shape, datatype message bytes, raw data bytes) of a compact dataset."""
        shape, dt, raw = (), None, None
        for t, p in self.messages(addr):
            if t == 1:  # dataspace
                ver, rank, fl = p[0], p[1], p[2]
                o = 4 if ver == 2 else 8
                shape = struct.unpack_from("<" + "Q" * rank, p, o) if rank else ()
            elif t == 3:
                dt = p
            elif t == 8:  # layout
                ver, cls = p[0], p[1]
                assert ver in (3, 4) and cls == 0, ("only compact layouts handled", ver, cls)
                sz = struct.unpack_from("<H", p, 2)[0]
                raw = p[4:4 + sz]
        return shape, dt, raw

    def numeric(self, addr):
        shape, dt, raw = self.dataset(addr)
        cls = dt[0] & 0x0F
        size = struct.unpack_from("<I", dt, 4)[0]
        if cls == 1 and size == 8:
            a = np.frombuffer(raw, dtype="<f8")
        elif cls == 0 and size == 8:
            a = np.frombuffer(raw, dtype="<i8")
        else:
            raise ValueError(("unhandled datatype", cls, size))
        return a.copy() if shape else a[0]

    def sparse(self, addr):
        shape, dt, raw = self.dataset(addr)
        # the compound datatype is a committed (shared) type under /_types; its layout is fixed:
        # {m::Int64, n::Int64, colptr::ref, rowval::ref, nzval::ref} = 40 bytes
        assert len(raw) == 40, (dt[0], len(raw))
        m, n, r_colptr, r_rowval, r_nzval = struct.unpack("<qqQQQ", raw)
        colptr = self.numeric(r_colptr)
        rowval = self.numeric(r_rowval)
        nzval = self.numeric(r_nzval)
        return m, n, colptr, rowval, nzval


def main():
    if not os.path.exists(SRC):
        sys.exit(f"{SRC} not found (this script only runs where the reference is mounted)")
    h = H5(SRC)
    top = h.links(h.root)
    print("datasets:", sorted(top))
    out = {}
    for name in ("q", "l", "u", "x_test", "y_test"):
        out[name] = np.asarray(h.numeric(top[name]), dtype=np.float64)
    out["obj_test"] = np.float64(h.numeric(top["obj_test"]))
    for name in ("P", "A"):
        m, n, colptr, rowval, nzval = h.sparse(top[name])
        out[name + "_shape"] = np.array([m, n], dtype=np.int64)
        out[name + "_colptr"] = colptr.astype(np.int64) - 1  # Julia is 1-based
        out[name + "_rowval"] = rowval.astype(np.int64) - 1
        out[name + "_nzval"] = nzval.astype(np.float64)
    np.savez(OUT, **out)
    for k, v in out.items():
        print(k, np.shape(v))
    print("obj_test =", repr(float(out["obj_test"])))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
