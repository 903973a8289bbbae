"""TEST INFRASTRUCTURE: a tiny HDF5 writer (superblock v0, version-1 object headers, symbol-table root group, contiguous
and chunked+deflate datasets) -- the "earliest" file layout h5py / h5netcdf produce -- to exercise
windgym_b200/hdf5_min.py without h5py.  Layout per the HDF5 File Format Specification 3.0."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, payload):
    payload = _pad8(payload)
    return struct.pack("<HHB3x", mtype, len(payload), 0) + payload


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        cls_bits = 0x11, (0x20, 0x3F if dt.itemsize == 8 else 0x1F, 0)          # class 1 v1; LE, mantissa norm, sign pos
        if dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        else:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    else:
        cls_bits = 0x10, (0x08 if dt.kind == "i" else 0, 0, 0)
        props = struct.pack("<HH", 0, 8 * dt.itemsize)
    return struct.pack("<BBBBI", cls_bits[0], *cls_bits[1], dt.itemsize) + props


def _dataspace_msg(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


def write(path, datasets, chunked=()):
    """datasets: {name: ndarray}; names in ``chunked`` are stored chunked (half-size chunks) with shuffle + deflate."""
    blob = bytearray(b"\0" * 2048)          # superblock + root structures are patched in at the end

    def put(b):
        off = len(blob)
        blob.extend(_pad8(b))
        return off

    entries = []
    for name in sorted(datasets):
        arr = np.ascontiguousarray(datasets[name])
        msgs = [_msg(0x01, _dataspace_msg(arr.shape)), _msg(0x03, _dtype_msg(arr.dtype))]
        if name in chunked:
            cshape = tuple(max(1, (s + 1) // 2) for s in arr.shape)
            esz = arr.dtype.itemsize
            keys = []
            import itertools
            for idx in itertools.product(*[range(0, s, c) for s, c in zip(arr.shape, cshape)]):
                blk = np.zeros(cshape, dtype=arr.dtype)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(idx, cshape, arr.shape))
                blk[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
                raw = np.frombuffer(blk.tobytes(), dtype=np.uint8).reshape(-1, esz).T.tobytes()      # shuffle
                comp = zlib.compress(raw, 4)
                keys.append((len(comp), idx, put(comp)))
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), UNDEF, UNDEF)
            for csize, idx, addr in keys:
                node += struct.pack("<II", csize, 0) + b"".join(struct.pack("<Q", o) for o in idx) + struct.pack("<Q", 0)
                node += struct.pack("<Q", addr)
            node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
            bt = put(node)
            msgs.append(_msg(0x0B, struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", esz, 0)
                             + struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 4, 0)))
            msgs.append(_msg(0x08, struct.pack("<BBB", 3, 2, arr.ndim + 1) + struct.pack("<Q", bt)
                             + b"".join(struct.pack("<I", c) for c in cshape) + struct.pack("<I", esz)))
        else:
            data = put(arr.tobytes())
            msgs.append(_msg(0x08, struct.pack("<BB", 3, 1) + struct.pack("<QQ", data, arr.nbytes)))
        entries.append((name, put(_object_header(msgs))))
    # local heap with the names, symbol node, B-tree, root object header
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for name, _ in entries:
        name_off[name] = len(heap_data)
        heap_data.extend(_pad8(name.encode() + b"\0"))
    hd = put(bytes(heap_data))
    heap = put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, hd))
    snod = b"SNOD" + struct.pack("<BxH", 1, len(entries))
    for name, addr in entries:
        snod += struct.pack("<QQII16x", name_off[name], addr, 0, 0)
    sn = put(snod)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<Q", 0) + struct.pack("<Q", sn) + \
        struct.pack("<Q", name_off[entries[-1][0]])
    bt = put(tree)
    root = put(_object_header([_msg(0x11, struct.pack("<QQ", bt, heap))]))
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBxHHI", 0, 0, 0, 0, 0, 8, 8, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(blob), UNDEF)
    sb += struct.pack("<QQII16x", 0, root, 0, 0)
    blob[:len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(blob))
