"""
Minimal pure-Python HDF5 reader -- just enough to open the `.h5ad` files Tangram's tests and tutorials use
(this image has no h5py / anndata).  Host-side I/O only; nothing here is on the hot path.

Supported: superblock v0/v1, v1 object headers (with continuation blocks), old-style groups (v1 B-tree + SNOD +
local heap), contiguous / compact / chunked (v1 chunk B-tree, no filters) datasets, fixed-point / IEEE float /
fixed-length string / variable-length string datatypes (global heap), attribute messages v1-v3.
Not supported (raises): compression filters, v2 object headers, new-style (fractal heap) groups.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class _Datatype:
    def __init__(self, cls, size, dtype=None, vlen_string=False, base=None):
        self.cls, self.size, self.dtype, self.vlen_string, self.base = cls, size, dtype, vlen_string, base


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        ver = b[8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} not supported")
        if b[13] != 8 or b[14] != 8:
            raise H5Error("only 8-byte offsets/lengths are supported")
        off = 24 if ver == 0 else 28
        self.base = struct.unpack_from("<Q", b, off)[0]
        root_entry = off + 32
        self.root_header = struct.unpack_from("<Q", b, root_entry + 8)[0]
        self._gcol = {}

    # ---------------------------------------------------------------- object headers
    def _messages(self, addr):
        b = self.buf
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} not supported")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                data = pos + 8
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", b, data)
                    blocks.append((caddr, clen))
                out.append((mtype, data, msize))
                pos = data + msize
        return out

    # ---------------------------------------------------------------- groups
    def _group_links(self, addr):
        links = {}
        for mtype, data, _ in self._messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", self.buf, data)
                heap_data = self._local_heap(heap)
                self._walk_group_btree(btree, heap_data, links)
        return links

    def _local_heap(self, addr):
        if self.buf[addr:addr + 4] != b"HEAP":
            raise H5Error("bad local heap")
        return struct.unpack_from("<Q", self.buf, addr + 24)[0]

    def _walk_group_btree(self, addr, heap_data, links):
        b = self.buf
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad B-tree node")
        _ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        pos = addr + 24
        for i in range(nent):
            child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0]
            if level > 0:
                self._walk_group_btree(child, heap_data, links)
            else:
                if b[child:child + 4] != b"SNOD":
                    raise H5Error("bad symbol node")
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                for s in range(nsym):
                    e = child + 8 + s * 40
                    name_off, hdr = struct.unpack_from("<QQ", b, e)
                    p = heap_data + name_off
                    name = b[p:b.index(b"\x00", p)].decode()
                    links[name] = hdr

    def is_group(self, addr):
        return any(m[0] == 0x11 for m in self._messages(addr))

    def resolve(self, path):
        addr = self.root_header
        for part in [p for p in path.split("/") if p]:
            links = self._group_links(addr)
            if part not in links:
                raise KeyError(path)
            addr = links[part]
        return addr

    def listdir(self, path="/"):
        return sorted(self._group_links(self.resolve(path)))

    def __contains__(self, path):
        try:
            self.resolve(path)
            return True
        except KeyError:
            return False

    # ---------------------------------------------------------------- datatypes / dataspaces
    def _datatype(self, pos):
        b = self.buf
        cv, b0, _b1, _b2, size = struct.unpack_from("<BBBBI", b, pos)
        cls = cv & 0x0F
        if cls == 0:
            signed = (b0 >> 3) & 1
            return _Datatype(0, size, np.dtype(("<" if not (b0 & 1) else ">") + ("i" if signed else "u") + str(size)))
        if cls == 1:
            return _Datatype(1, size, np.dtype(("<" if not (b0 & 1) else ">") + "f" + str(size)))
        if cls == 3:
            return _Datatype(3, size)
        if cls == 9:
            base = self._datatype(pos + 8)
            return _Datatype(9, size, vlen_string=(b0 & 0x0F) == 1, base=base)
        if cls == 7:   # object reference
            return _Datatype(7, size, np.dtype("<u8"))
        if cls == 8:   # enum (anndata stores booleans this way): read the base integers
            base = self._datatype(pos + 8)
            return _Datatype(0, size, base.dtype)
        raise H5Error(f"datatype class {cls} not supported")

    def _dataspace(self, pos):
        b = self.buf
        ver, rank, flags = struct.unpack_from("<BBB", b, pos)
        if ver == 1:
            dpos = pos + 8
        elif ver == 2:
            dpos = pos + 4
        else:
            raise H5Error(f"dataspace version {ver}")
        return tuple(struct.unpack_from("<" + "Q" * rank, b, dpos)) if rank else ()

    # ---------------------------------------------------------------- global heap (vlen strings)
    def _gcol_objects(self, addr):
        if addr in self._gcol:
            return self._gcol[addr]
        b = self.buf
        if b[addr:addr + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        size = struct.unpack_from("<Q", b, addr + 8)[0]
        pos, end, objs = addr + 16, addr + size, {}
        while pos + 16 <= end:
            idx, _rc, _r, osz = struct.unpack_from("<HHIQ", b, pos)
            if idx == 0:
                break
            objs[idx] = (pos + 16, osz)
            pos += 16 + ((osz + 7) // 8) * 8
        self._gcol[addr] = objs
        return objs

    def _decode(self, raw, dt, shape):
        n = int(np.prod(shape)) if shape else 1
        if dt.cls in (0, 1, 7):
            return np.frombuffer(raw, dtype=dt.dtype, count=n).reshape(shape).copy()
        if dt.cls == 3:
            out = [raw[i * dt.size:(i + 1) * dt.size].split(b"\x00")[0].decode("utf-8", "replace") for i in range(n)]
            return np.array(out, dtype=object).reshape(shape)
        if dt.cls == 9:
            out = []
            for i in range(n):
                length, gaddr, gidx = struct.unpack_from("<IQI", raw, i * 16)
                if length == 0 or gaddr in (0, UNDEF):
                    out.append("" if dt.vlen_string else np.array([]))
                    continue
                p, _sz = self._gcol_objects(gaddr)[gidx]
                if dt.vlen_string:
                    out.append(self.buf[p:p + length].decode("utf-8", "replace"))
                else:
                    out.append(np.frombuffer(self.buf, dtype=dt.base.dtype, count=length, offset=p).copy())
            return np.array(out, dtype=object).reshape(shape)
        raise H5Error("cannot decode")

    # ---------------------------------------------------------------- datasets
    def read(self, path):
        addr = self.resolve(path)
        dt = shape = layout = None
        for mtype, data, _ in self._messages(addr):
            if mtype == 0x01:
                shape = self._dataspace(data)
            elif mtype == 0x03:
                dt = self._datatype(data)
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                raise H5Error(f"{path}: filtered (compressed) datasets are not supported")
        if dt is None or shape is None or layout is None:
            raise H5Error(f"{path} is not a dataset")
        b = self.buf
        ver, lclass = struct.unpack_from("<BB", b, layout)
        if ver != 3:
            raise H5Error(f"data layout version {ver} not supported")
        n = int(np.prod(shape)) if shape else 1
        esize = dt.size
        if lclass == 0:
            size = struct.unpack_from("<H", b, layout + 2)[0]
            raw = b[layout + 4:layout + 4 + size]
        elif lclass == 1:
            daddr, dsize = struct.unpack_from("<QQ", b, layout + 2)
            raw = b"" if daddr == UNDEF else b[daddr:daddr + n * esize]
            if daddr == UNDEF:
                raw = bytes(n * esize)
        elif lclass == 2:
            nd = b[layout + 2]
            btree = struct.unpack_from("<Q", b, layout + 3)[0]
            cdims = struct.unpack_from("<" + "I" * nd, b, layout + 11)
            raw = self._read_chunked(btree, shape, cdims[:-1], esize)
        else:
            raise H5Error("unknown layout class")
        return self._decode(raw, dt, shape)

    def _read_chunked(self, btree, shape, cshape, esize):
        out = np.zeros(shape, dtype=np.dtype(("V", esize)))
        rank = len(shape)
        if btree == UNDEF:
            return out.tobytes()

        def walk(addr):
            b = self.buf
            if b[addr:addr + 4] != b"TREE":
                raise H5Error("bad chunk B-tree node")
            _ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
            keysz = 8 + 8 * (rank + 1)
            pos = addr + 24
            for i in range(nent):
                k = pos + i * (keysz + 8)
                csize, fmask = struct.unpack_from("<II", b, k)
                offs = struct.unpack_from("<" + "Q" * rank, b, k + 8)
                child = struct.unpack_from("<Q", b, k + keysz)[0]
                if level > 0:
                    walk(child)
                    continue
                if fmask != 0:
                    raise H5Error("filtered chunk")
                chunk = np.frombuffer(b, dtype=out.dtype, count=int(np.prod(cshape)), offset=child).reshape(cshape)
                sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                sl_in = tuple(slice(0, so.stop - so.start) for so in sl_out)
                out[sl_out] = chunk[sl_in]
        walk(btree)
        return out.tobytes()

    # ---------------------------------------------------------------- attributes
    def attrs(self, path):
        addr = self.resolve(path)
        b = self.buf
        out = {}
        for mtype, data, _ in self._messages(addr):
            if mtype != 0x0C:
                continue
            ver = b[data]
            if ver == 1:
                nsz, dsz, ssz = struct.unpack_from("<HHH", b, data + 2)
                p = data + 8
                pad = lambda x: (x + 7) // 8 * 8   # noqa: E731
            elif ver in (2, 3):
                nsz, dsz, ssz = struct.unpack_from("<HHH", b, data + 2)
                p = data + 8 + (1 if ver == 3 else 0)
                pad = lambda x: x   # noqa: E731
            else:
                continue
            name = b[p:p + nsz].split(b"\x00")[0].decode()
            p += pad(nsz)
            try:
                dt = self._datatype(p)
                shape = self._dataspace(p + pad(dsz))
                p2 = p + pad(dsz) + pad(ssz)
                n = int(np.prod(shape)) if shape else 1
                val = self._decode(b[p2:p2 + n * dt.size], dt, shape)
                out[name] = val.item() if val.shape == () else val
            except H5Error:
                out[name] = None
        return out
