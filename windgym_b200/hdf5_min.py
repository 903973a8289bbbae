"""Minimal pure-Python HDF5 reader for NetCDF-4 turbulence boxes (no h5py / netCDF4 in the GPU image).

The reference loads its Mann boxes with ``MannTurbulenceField.from_netcdf`` (``Wind_Farm_Env.py:611-618``,
``FarmEval.update_tf`` ``FarmEval.py:86-90``); hipersim / xarray write them as NetCDF-4, i.e. HDF5 files holding a
handful of fixed-size numeric datasets (``uvw`` [3, Nx, Ny, Nz] and the coordinate axes).  This module reads exactly
that subset of the HDF5 file format (HDF5 File Format Specification 3.0):

* superblock versions 0-3; object headers version 1 and 2 (with continuation blocks);
* groups with compact links (link messages) or dense links (fractal heap, direct blocks scanned sequentially) and
  old-style symbol-table groups (B-tree v1 + local heap);
* datasets with CONTIGUOUS or COMPACT layout, and CHUNKED layout (B-tree v1 chunk index) without filters or with the
  deflate (+ shuffle) filters netCDF4 applies when ``zlib=True``;
* fixed-point and floating-point little/big-endian element types.

Anything else (variable-length strings, compound types, virtual datasets, other filters ...) raises
``NotImplementedError`` naming ``scripts/convert_mann_netcdf.py`` as the way out.  Validated against the reference's own
NetCDF-4 file ``examples/PPO_eval.nc`` (tests/test_host_logic.py, build container only) and a synthetic writer.
"""
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(NotImplementedError):
    pass


def _unsupported(what):
    raise Hdf5Error(f"hdf5_min: {what} is not supported by the built-in reader; convert the file to .npz with "
                    "scripts/convert_mann_netcdf.py (needs netCDF4, h5py or xarray)")


class _Reader:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        self.base = 0
        self._superblock()

    # ------------------------------------------------------------------------------------------ primitives
    def u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _superblock(self):
        b = self.buf
        sig = b"\x89HDF\r\n\x1a\n"
        off = 0
        while b[off:off + 8] != sig:       # the superblock may sit at 0, 512, 1024, ...
            off = 512 if off == 0 else off * 2
            if off >= len(b):
                raise ValueError("not an HDF5 file (signature not found)")
        self.sb_off = off
        ver = b[off + 8]
        if ver in (0, 1):
            self.so, self.sl = b[off + 13], b[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = self.u(p, self.so)
            p += 4 * self.so                # base, free-space, end-of-file, driver-info addresses
            # root group symbol table entry: link name offset, object header address, cache type, ...
            self.root = self.u(p + self.so, self.so)
        elif ver in (2, 3):
            self.so, self.sl = b[off + 9], b[off + 10]
            p = off + 12
            self.base = self.u(p, self.so)
            self.root = self.u(p + 3 * self.so, self.so)
        else:
            _unsupported(f"superblock version {ver}")

    # ------------------------------------------------------------------------------------------ object headers
    def messages(self, addr):
        """List of (type, flags, payload offset, payload size) of the object header at ``addr``."""
        b = self.buf
        addr += self.base
        out = []
        if b[addr:addr + 4] == b"OHDR":                       # version 2
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16                                       # access / modification / change / birth times
            if flags & 0x10:
                p += 4                                        # max compact / min dense attributes
            nsz = 1 << (flags & 3)
            chunk0 = self.u(p, nsz)
            p += nsz
            track = bool(flags & 0x04)
            blocks = [(p, chunk0)]
            while blocks:
                start, size = blocks.pop(0)
                q, end = start, start + size
                while q + 4 <= end:
                    mtype, msize, mflags = b[q], self.u(q + 1, 2), b[q + 3]
                    q += 4 + (2 if track else 0)
                    if q + msize > end:
                        break
                    if mtype == 0x10:                         # continuation: OCHK block (signature + messages + checksum)
                        caddr, clen = self.u(q, self.so) + self.base, self.u(q + self.so, self.sl)
                        if b[caddr:caddr + 4] != b"OCHK":
                            raise ValueError("corrupt object header continuation")
                        blocks.append((caddr + 4, clen - 8))
                    elif mtype != 0:
                        out.append((mtype, mflags, q, msize))
                    q += msize
        else:                                                 # version 1
            if b[addr] != 1:
                raise ValueError(f"unknown object header version {b[addr]} at {addr}")
            nmsg, hsize = self.u(addr + 2, 2), self.u(addr + 8, 4)
            blocks = [(addr + 16, hsize)]
            while blocks and len(out) < 4096:
                start, size = blocks.pop(0)
                q, end = start, start + size
                while q + 8 <= end:
                    mtype, msize, mflags = self.u(q, 2), self.u(q + 2, 2), b[q + 4]
                    q += 8
                    if mtype == 0x10:
                        blocks.append((self.u(q, self.so) + self.base, self.u(q + self.so, self.sl)))
                    elif mtype != 0:
                        out.append((mtype, mflags, q, msize))
                    q += msize
        return out

    # ------------------------------------------------------------------------------------------ groups
    def _link_message(self, q):
        """Parse one link message at ``q``: returns (name, object header address or None, next offset)."""
        b = self.buf
        ver, flags = b[q], b[q + 1]
        if ver != 1:
            raise ValueError("not a link message")
        p = q + 2
        ltype = 0
        if flags & 0x08:
            ltype = b[p]; p += 1
        if flags & 0x04:
            p += 8                                            # creation order
        if flags & 0x10:
            p += 1                                            # character set
        nsz = 1 << (flags & 3)
        nlen = self.u(p, nsz); p += nsz
        name = b[p:p + nlen].decode("utf-8", "replace"); p += nlen
        if ltype == 0:                                        # hard link
            return name, self.u(p, self.so), p + self.so
        if ltype == 1:                                        # soft link: length + path
            return name, None, p + 2 + self.u(p, 2)
        return name, None, p

    def links(self, addr):
        """{name: object header address} of the group at ``addr``."""
        b = self.buf
        out = {}
        for mtype, _, q, size in self.messages(addr):
            if mtype == 0x06:                                 # link message (compact storage)
                name, target, _ = self._link_message(q)
                if target is not None:
                    out[name] = target
            elif mtype == 0x02:                               # link info: dense storage in a fractal heap
                flags = b[q + 1]
                p = q + 2 + (8 if flags & 1 else 0)
                heap = self.u(p, self.so)
                if heap != UNDEF & ((1 << (8 * self.so)) - 1):
                    out.update(self._dense_links(heap + self.base))
            elif mtype == 0x11:                               # symbol table (old-style group)
                out.update(self._symbol_table(self.u(q, self.so) + self.base, self.u(q + self.so, self.so) + self.base))
        return out

    def _dense_links(self, frhp):
        """Links stored as managed objects of a fractal heap.  The heap's direct blocks hold the link messages back to
        back; they are scanned sequentially (the name index B-tree is not needed to enumerate them)."""
        b = self.buf
        if b[frhp:frhp + 4] != b"FRHP":
            raise ValueError("fractal heap header expected")
        p = frhp + 5
        p += 2 + 2 + 1                                        # heap id length, I/O filter length, flags
        p += 4                                                # max size of managed objects
        p += self.sl + self.so                                # next huge id, huge-object B-tree address
        p += self.sl + self.so                                # free space in managed blocks, free-space manager address
        p += 4 * self.sl                                      # managed space, allocated space, iterator offset, n managed objects
        p += 4 * self.sl                                      # huge size/count, tiny size/count
        table_width = self.u(p, 2); p += 2
        start_block = self.u(p, self.sl); p += self.sl
        max_direct = self.u(p, self.sl); p += self.sl
        max_heap_bits = self.u(p, 2); p += 2
        p += 2                                                # starting # of rows in the root indirect block
        root = self.u(p, self.so); p += self.so
        cur_rows = self.u(p, 2)
        out = {}
        off_bytes = (max_heap_bits + 7) // 8

        def scan_direct(addr, size):
            addr += self.base
            if b[addr:addr + 4] != b"FHDB":
                return
            q = addr + 5 + self.so + off_bytes                # signature, version, heap header address, block offset
            flags = self.u(frhp + 5 + 4, 1)
            if flags & 2:
                q += 4                                        # checksum of the block
            end = addr + size
            while q + 10 < end and b[q] == 1:                 # link messages, version 1
                try:
                    name, target, nxt = self._link_message(q)
                except Exception:
                    break
                if not name or nxt <= q:
                    break
                if target is not None:
                    out[name] = target
                q = nxt

        if cur_rows == 0:
            scan_direct(root, start_block)
        else:                                                 # root indirect block: rows of direct blocks (doubling table)
            ib = root + self.base
            if b[ib:ib + 4] != b"FHIB":
                raise ValueError("fractal heap indirect block expected")
            q = ib + 5 + self.so + off_bytes
            size = start_block
            for row in range(cur_rows):
                if row >= 2:
                    size *= 2
                if size > max_direct:
                    _unsupported("a fractal heap with nested indirect blocks (very large groups)")
                for _ in range(table_width):
                    child = self.u(q, self.so); q += self.so
                    if child != UNDEF & ((1 << (8 * self.so)) - 1):
                        scan_direct(child, size)
        return out

    def _symbol_table(self, btree, heap):
        b = self.buf
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("local heap expected")
        data = self.u(heap + 8 + 2 * self.sl, self.so) + self.base
        out = {}

        def node(addr):
            if b[addr:addr + 4] == b"TREE":
                level, n = b[addr + 5], self.u(addr + 6, 2)
                p = addr + 8 + 2 * self.so
                for i in range(n):
                    p += self.sl                              # key
                    child = self.u(p, self.so) + self.base; p += self.so
                    node(child)
            elif b[addr:addr + 4] == b"SNOD":
                n = self.u(addr + 6, 2)
                p = addr + 8
                for i in range(n):
                    noff, ohdr = self.u(p, self.so), self.u(p + self.so, self.so)
                    e = b.index(b"\0", data + noff)
                    out[b[data + noff:e].decode()] = ohdr
                    p += 2 * self.so + 4 + 4 + 16
        node(btree)
        return out

    # ------------------------------------------------------------------------------------------ datasets
    def dataset(self, addr):
        """numpy array of the dataset whose object header is at ``addr`` (None for non-numeric element types)."""
        b = self.buf
        shape, dtype, layout, filters = None, None, None, []
        for mtype, _, q, size in self.messages(addr):
            if mtype == 0x01:                                 # dataspace
                ver, rank, flags = b[q], b[q + 1], b[q + 2]
                p = q + (8 if ver == 1 else 4)
                shape = tuple(self.u(p + i * self.sl, self.sl) for i in range(rank))
            elif mtype == 0x03:                               # datatype
                cls, bits0 = b[q] & 0x0F, b[q + 1]
                sz = self.u(q + 4, 4)
                order = ">" if bits0 & 1 else "<"
                if cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = np.dtype(f"{order}{'i' if signed else 'u'}{sz}")
                elif cls == 1:
                    dtype = np.dtype(f"{order}f{sz}")
                else:
                    dtype = None                              # strings, compounds, references ...: not a numeric array
            elif mtype == 0x08:                               # data layout
                ver = b[q]
                if ver == 3:
                    cls = b[q + 1]
                    if cls == 0:                              # compact: data inside the message
                        n = self.u(q + 2, 2)
                        layout = ("compact", q + 4, n)
                    elif cls == 1:
                        layout = ("contiguous", self.u(q + 2, self.so), self.u(q + 2 + self.so, self.sl))
                    elif cls == 2:
                        rank = b[q + 2]
                        bt = self.u(q + 3, self.so)
                        dims = tuple(self.u(q + 3 + self.so + 4 * i, 4) for i in range(rank))
                        layout = ("chunked", bt, dims)
                    else:
                        _unsupported(f"data layout class {cls}")
                elif ver == 4:
                    cls = b[q + 1]
                    if cls == 1:
                        layout = ("contiguous", self.u(q + 2, self.so), self.u(q + 2 + self.so, self.sl))
                    elif cls == 0:
                        layout = ("compact", q + 4, self.u(q + 2, 2))
                    else:
                        _unsupported("version-4 chunked / virtual data layout")
                else:
                    _unsupported(f"data layout message version {ver}")
            elif mtype == 0x0B:                               # filter pipeline
                ver, nf = b[q], b[q + 1]
                p = q + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self.u(p, 2)
                    if ver == 1 or fid >= 256:
                        nlen = self.u(p + 2, 2); ncd = self.u(p + 6, 2); p += 8
                        p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    else:
                        ncd = self.u(p + 4, 2); p += 6
                    cd = [self.u(p + 4 * i, 4) for i in range(ncd)]
                    p += 4 * ncd + (4 if (ver == 1 and ncd % 2) else 0)
                    filters.append((fid, cd))
        if shape is None or layout is None:
            return None
        if dtype is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            raw = b[layout[1]:layout[1] + layout[2]]
            return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()
        if layout[0] == "contiguous":
            a = layout[1]
            if a == UNDEF & ((1 << (8 * self.so)) - 1):       # never written: fill value (zeros)
                return np.zeros(shape, dtype=dtype.newbyteorder("="))
            a += self.base
            return np.frombuffer(b, dtype=dtype, count=n, offset=a).reshape(shape).copy()
        return self._chunked(layout[1] + self.base, layout[2], shape, dtype, filters)

    def _chunked(self, btree, cdims, shape, dtype, filters):
        b = self.buf
        for fid, _ in filters:
            if fid not in (1, 2):                             # 1 deflate, 2 shuffle
                _unsupported(f"HDF5 filter id {fid}")
        rank = len(shape)
        chunk = cdims[:rank]
        out = np.zeros(shape, dtype=dtype)
        esz = dtype.itemsize

        def node(addr):
            if b[addr:addr + 4] != b"TREE":
                raise ValueError("chunk B-tree node expected")
            level, n = b[addr + 5], self.u(addr + 6, 2)
            p = addr + 8 + 2 * self.so
            for i in range(n):
                csize, fmask = self.u(p, 4), self.u(p + 4, 4)
                offs = tuple(self.u(p + 8 + 8 * k, 8) for k in range(rank))
                p += 8 + 8 * (rank + 1)
                child = self.u(p, self.so) + self.base; p += self.so
                if level > 0:
                    node(child)
                    continue
                raw = b[child:child + csize]
                for fid, cd in reversed(filters):
                    idx = [f for f, _ in filters].index(fid)
                    if fmask & (1 << idx):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        a = np.frombuffer(raw, dtype=np.uint8)
                        m = a.size // esz
                        raw = a[:m * esz].reshape(esz, m).T.tobytes() + a[m * esz:].tobytes()
                blk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk))).reshape(chunk)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        node(btree)
        return out


def read_datasets(path, names=None):
    """{name: ndarray} of the numeric datasets in the root group of an HDF5 / NetCDF-4 file (``names``: only those)."""
    r = _Reader(path)
    out = {}
    for name, addr in r.links(r.root).items():
        if names is not None and name not in names:
            continue
        try:
            arr = r.dataset(addr)
        except Hdf5Error:
            if names is not None:
                raise
            continue
        if arr is not None:
            out[name] = arr.astype(arr.dtype.newbyteorder("=")) if arr.dtype.byteorder == ">" else arr
    return out


def list_datasets(path):
    r = _Reader(path)
    return sorted(r.links(r.root))


def read_mann_box(path):
    """(uvw [3, Nx, Ny, Nz] float32, (dx, dy, dz)) of a hipersim / dynamiks ``MannTurbulenceField.to_netcdf`` file: the
    4-D dataset with a leading axis of 3 (named ``uvw``, or ``__xarray_dataarray_variable__`` when xarray wrote an
    unnamed DataArray) and the 1-D coordinate axes ``x``, ``y``, ``z``."""
    ds = read_datasets(path)
    cand = [k for k, v in ds.items() if v.ndim == 4 and v.shape[0] == 3]
    if "uvw" in cand:
        cand = ["uvw"]
    if len(cand) != 1 or any(a not in ds or ds[a].ndim != 1 for a in ("x", "y", "z")):
        raise ValueError(f"{path}: expected one [3, Nx, Ny, Nz] dataset and 1-D axes x, y, z; the root group holds "
                         f"{ {k: v.shape for k, v in ds.items()} }")
    uvw = ds[cand[0]]
    if uvw.shape[1:] != (ds["x"].size, ds["y"].size, ds["z"].size):
        raise ValueError(f"{path}: field {uvw.shape} does not match the axes ({ds['x'].size}, {ds['y'].size}, {ds['z'].size})")
    dxyz = tuple(float(ds[a][1] - ds[a][0]) for a in ("x", "y", "z"))
    return np.ascontiguousarray(uvw, dtype=np.float32), dxyz
