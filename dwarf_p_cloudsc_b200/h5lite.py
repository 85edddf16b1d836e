"""Dependency-free reader/writer for the small HDF5 subset the CLOUDSC dwarf uses.

The dwarf's ``input.h5`` / ``reference.h5`` (reference: serialbox2hdf5/serialbox2hdf5.py:25-56,
read by src/cloudsc_c/cloudsc/load_state.c:26-41) are flat files: one root group holding
contiguous, unfiltered, little-endian datasets of float64 / int32 / int64 / uint8.  Neither
libhdf5 nor h5py exists in this image, so this module implements exactly that subset of the
HDF5 file format (superblock v0, v1 object headers, group B-tree v1 + SNOD + local heap,
dataspace v1/v2, datatype classes 0/1, layout v3 contiguous/compact):

* ``read_h5(path)``  -> ``{name: ndarray}`` for every dataset in the root group.
* ``write_h5(path, {name: ndarray})`` -> a file in the same layout libhdf5 1.8/1.10 writes
  with default (earliest) format, so the stock HDF5 loader of the reference can read it.

The C twin of the reader is ``host/h5lite.c`` (used by ``load_state``); tests check both
against each other and against ``reference.h5`` as shipped by the reference.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF

_MSG_DATASPACE = 0x0001
_MSG_DATATYPE = 0x0003
_MSG_FILL_OLD = 0x0004
_MSG_FILL = 0x0005
_MSG_LAYOUT = 0x0008
_MSG_CONT = 0x0010
_MSG_SYMTAB = 0x0011
_MSG_MTIME = 0x0012


# --------------------------------------------------------------------------------------
# reader
# --------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file (bad signature)")
        ver = buf[8]
        if ver != 0:
            raise ValueError(f"unsupported superblock version {ver} (only v0)")
        if buf[13] != 8 or buf[14] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        self.base = struct.unpack_from("<Q", buf, 24)[0]
        # root symbol-table entry at 56
        self.root_ohdr = struct.unpack_from("<Q", buf, 56 + 8)[0]

    def _messages(self, addr: int):
        """Yield (type, payload-bytes) of a version-1 object header, following continuations."""
        b = self.b
        ver, _, nmsg, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise ValueError(f"unsupported object header version {ver}")
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                seen += 1
                if mtype == _MSG_CONT:
                    off, length = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.base + off, length))
                else:
                    yield mtype, body

    def _group_entries(self, btree: int, heap: int) -> Dict[str, int]:
        b = self.b
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0] + self.base
        out: Dict[str, int] = {}

        def name_at(off: int) -> str:
            s = heap_data + off
            e = b.index(b"\0", s)
            return b[s:e].decode("ascii")

        def walk(node: int):
            if b[node:node + 4] == b"TREE":
                ntype, _level, nent = struct.unpack_from("<BBH", b, node + 4)
                if ntype != 0:
                    raise ValueError("not a group B-tree")
                pos = node + 24
                for i in range(nent):
                    child = struct.unpack_from("<Q", b, pos + 8 + 16 * i)[0]
                    walk(self.base + child)
            elif b[node:node + 4] == b"SNOD":
                nsym = struct.unpack_from("<H", b, node + 6)[0]
                for i in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", b, node + 8 + 40 * i)
                    out[name_at(noff)] = self.base + ohdr
            else:
                raise ValueError("unknown group node signature")

        walk(btree)
        return out

    def root(self) -> Dict[str, int]:
        for mtype, body in self._messages(self.root_ohdr):
            if mtype == _MSG_SYMTAB:
                bt, hp = struct.unpack_from("<QQ", body, 0)
                return self._group_entries(self.base + bt, self.base + hp)
        raise ValueError("root group has no symbol table message")

    def dataset(self, addr: int) -> np.ndarray:
        dims = None
        dtype = None
        raw = None
        for mtype, body in self._messages(addr):
            if mtype == _MSG_DATASPACE:
                ver, rank = body[0], body[1]
                off = 8 if ver == 1 else 4
                dims = struct.unpack_from("<%dQ" % rank, body, off) if rank else ()
            elif mtype == _MSG_DATATYPE:
                cls = body[0] & 0x0F
                bits0 = body[1]
                size = struct.unpack_from("<I", body, 4)[0]
                if bits0 & 1:
                    raise ValueError("big-endian datasets not supported")
                if cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = np.dtype("<%s%d" % ("i" if signed else "u", size))
                elif cls == 1:
                    dtype = np.dtype("<f%d" % size)
                else:
                    raise ValueError(f"unsupported datatype class {cls}")
            elif mtype == _MSG_LAYOUT:
                ver, lcls = body[0], body[1]
                if ver != 3:
                    raise ValueError(f"unsupported layout version {ver}")
                if lcls == 1:
                    a, n = struct.unpack_from("<QQ", body, 2)
                    raw = b"" if a == _UNDEF else self.b[self.base + a:self.base + a + n]
                elif lcls == 0:
                    n = struct.unpack_from("<H", body, 2)[0]
                    raw = body[4:4 + n]
                else:
                    raise ValueError("chunked layout not supported")
        if dims is None or dtype is None or raw is None:
            raise ValueError("object is not a simple dataset")
        count = int(np.prod(dims)) if dims else 1
        if len(raw) < count * dtype.itemsize:
            # never-written dataset: HDF5 semantics = fill value (zero)
            return np.zeros(dims, dtype=dtype)
        return np.frombuffer(raw, dtype=dtype, count=count).reshape(dims).copy()


def read_h5(path: str) -> Dict[str, np.ndarray]:
    """Read every root-level dataset of ``path`` (see module docstring for the subset)."""
    with open(path, "rb") as f:
        rd = _Reader(f.read())
    return {name: rd.dataset(addr) for name, addr in rd.root().items()}


# --------------------------------------------------------------------------------------
# writer
# --------------------------------------------------------------------------------------
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _datatype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        # IEEE little-endian float: class 1, version 1
        if dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)  # mantissa normalisation = implied msb; sign bit position 63
        elif dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)
        else:
            raise ValueError(dt)
        return struct.pack("<BBBBI", 0x11, bits[0], bits[1], bits[2], dt.itemsize) + props
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        props = struct.pack("<HH", 0, 8 * dt.itemsize)
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + props
    raise ValueError(f"unsupported dtype {dt}")


def _dataspace_msg(shape) -> bytes:
    rank = len(shape)
    # version 1, flags bit0 = max dims present (h5py/libhdf5 always writes them)
    return struct.pack("<BBB5x", 1, rank, 1) + struct.pack("<%dQ" % rank, *shape) * 2


def write_h5(path: str, datasets: Dict[str, np.ndarray]) -> None:
    """Write ``datasets`` as contiguous root-level datasets (earliest-format HDF5)."""
    names = sorted(datasets)  # group B-tree order = strcmp order
    for n in names:
        if not n or "/" in n:
            raise ValueError(f"bad dataset name {n!r}")
    arrays = {}
    for n in names:
        a = np.ascontiguousarray(datasets[n])
        if a.dtype.kind == "b":
            a = a.astype(np.uint8)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        arrays[n] = a

    LEAF_K = 4           # up to 2K = 8 symbols per SNOD (libhdf5 default)
    PER_NODE = 2 * LEAF_K
    groups = [names[i:i + PER_NODE] for i in range(0, len(names), PER_NODE)] or [[]]
    INT_K = max(16, (len(groups) + 1) // 2)  # one B-tree node holds every SNOD
    if INT_K > 0x7FFF:
        raise ValueError("too many datasets")

    # ---- local heap data: "" at 0, then names
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode("ascii") + b"\0")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)  # one free block: next=1 (none), size 16
    heap_size = len(heap)

    # ---- file layout
    pos = 96                                   # after superblock
    root_ohdr = pos; pos += 16 + 24            # header + one symbol-table message
    btree = pos; pos += 24 + (2 * INT_K + 1) * 8 + 2 * INT_K * 8
    heap_hdr = pos; pos += 32
    heap_data = pos; pos += heap_size
    snod_size = 8 + PER_NODE * 40
    snods = []
    for _ in groups:
        snods.append(pos); pos += snod_size
    ohdr_size = {}
    ohdr_addr = {}
    for n in names:
        a = arrays[n]
        msgs = (_msg(_MSG_DATASPACE, _dataspace_msg(a.shape)) +
                _msg(_MSG_DATATYPE, _datatype_msg(a.dtype), flags=1) +
                _msg(_MSG_FILL, struct.pack("<BBBBI", 2, 2, 2, 1, 0)) +
                _msg(_MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, 0, 0)))
        ohdr_addr[n] = pos
        ohdr_size[n] = len(msgs)
        pos += 16 + len(msgs)
    pos += -pos % 8
    data_addr = {}
    for n in names:
        data_addr[n] = pos
        pos += arrays[n].nbytes
        pos += -pos % 8
    eof = pos

    out = bytearray(eof)
    # superblock v0
    out[0:8] = _SIG
    struct.pack_into("<BBBBBBBBHHI", out, 8, 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INT_K, 0)
    struct.pack_into("<QQQQ", out, 24, 0, _UNDEF, eof, _UNDEF)
    struct.pack_into("<QQII", out, 56, 0, root_ohdr, 1, 0)
    struct.pack_into("<QQ", out, 80, btree, heap_hdr)
    # root object header
    struct.pack_into("<BBHII", out, root_ohdr, 1, 0, 1, 1, 24)
    out[root_ohdr + 16:root_ohdr + 40] = _msg(_MSG_SYMTAB, struct.pack("<QQ", btree, heap_hdr))
    # B-tree root node (level 0 -> children are SNODs)
    out[btree:btree + 4] = b"TREE"
    struct.pack_into("<BBHQQ", out, btree + 4, 0, 0, len(groups) if names else 0, _UNDEF, _UNDEF)
    p = btree + 24
    struct.pack_into("<Q", out, p, 0)  # key 0 = "" (heap offset 0)
    p += 8
    if names:
        for g, saddr in zip(groups, snods):
            struct.pack_into("<QQ", out, p, saddr, name_off[g[-1]])
            p += 16
    # local heap
    out[heap_hdr:heap_hdr + 4] = b"HEAP"
    struct.pack_into("<B3xQQQ", out, heap_hdr + 4, 0, heap_size, free_off, heap_data)
    out[heap_data:heap_data + heap_size] = heap
    # symbol nodes
    for g, saddr in zip(groups, snods):
        out[saddr:saddr + 4] = b"SNOD"
        struct.pack_into("<BBH", out, saddr + 4, 1, 0, len(g))
        for i, n in enumerate(g):
            struct.pack_into("<QQII", out, saddr + 8 + 40 * i, name_off[n], ohdr_addr[n], 0, 0)
    # datasets
    for n in names:
        a = arrays[n]
        oa = ohdr_addr[n]
        struct.pack_into("<BBHII", out, oa, 1, 0, 4, 1, ohdr_size[n])
        msgs = (_msg(_MSG_DATASPACE, _dataspace_msg(a.shape)) +
                _msg(_MSG_DATATYPE, _datatype_msg(a.dtype), flags=1) +
                _msg(_MSG_FILL, struct.pack("<BBBBI", 2, 2, 2, 1, 0)) +
                _msg(_MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, data_addr[n], a.nbytes)))
        out[oa + 16:oa + 16 + len(msgs)] = msgs
        out[data_addr[n]:data_addr[n] + a.nbytes] = a.tobytes()
    with open(path, "wb") as f:
        f.write(bytes(out))
