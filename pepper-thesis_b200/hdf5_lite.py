"""A small HDF5 writer / reader for the on-disk hand-off of the reference: the classic file layout
every HDF5 library reads -- superblock version 0, version-1 object headers, groups as symbol tables (version-1 B-tree +
local heap + symbol-table nodes), contiguous datasets -- for exactly the types the reference's files hold
(/root/reference/pepper_variant/modules/python/DataStore.py:54-71, DataStorePredict.py:49-66): int8 / uint8 / int32 / int64 /
float64 arrays, fixed-length byte strings (numpy 'S'), variable-length strings (h5py.special_dtype(vlen=str): global heap
collections) and scalar strings (the yaml meta entries).

h5py / libhdf5 are absent from this image, so nothing here can be checked against the library itself. What pins it: the
READER below parses a file written by the real library -- scipy ships one (scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat,
written by the HDF5 library inside MATLAB: same superblock, symbol-table group, object-header, dataspace, datatype and layout messages) -- and
the WRITER's files are read back by that same reader and checked structure by structure (tests/test_hdf5_lite.py). The
format follows the HDF5 File Format Specification version 2.0 (sections III.A-III.E, IV.A.1-2).

Host-side I/O only; no device code involved."""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
GROUP_LEAF_K = 4            # a symbol-table node holds up to 2K entries
GROUP_INTERNAL_K = 16       # a B-tree node holds up to 2K children
GCOL_MIN = 4096

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT = 0x0, 0x1, 0x3, 0x4, 0x5, 0x8
MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE, MSG_MTIME = 0xC, 0x10, 0x11, 0x12


class VlenStr:
    """marks a dataset of variable-length strings (what h5py.special_dtype(vlen=str) makes)"""

    def __init__(self, values, shape=None):
        arr = np.asarray(values, dtype=object)
        self.shape = tuple(arr.shape if shape is None else shape)
        self.values = [v if isinstance(v, bytes) else str(v).encode("utf-8") for v in arr.reshape(-1)]


def _pad8(n: int) -> int:
    return (n + 7) & ~7


# ---------------------------------------------------------------------------------------------------------------------
# datatype / dataspace messages
# ---------------------------------------------------------------------------------------------------------------------
def _datatype_message(dt: Union[np.dtype, str]) -> bytes:
    if isinstance(dt, str) and dt == "vlen_str":
        # class 9 (variable length), version 1; bits: type 1 = string, padding 0 (null terminated), character set 1 (UTF-8:
        # h5py's str). The base type is the library's H5T_C_S1 as it stands (1 byte, null terminated, ASCII): setting the
        # size of a string type to "variable" wraps that type, the character set lives on the variable-length level.
        base = struct.pack("<BBBBI", 0x13, 0x00, 0x00, 0x00, 1)
        return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0x00, 16) + base
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00                      # bit 3: signed; byte order little endian, no padding
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize == 8:
        # IEEE double, little endian: mantissa normalisation 2 (implied), sign bit 63; exponent 52..62, bias 1023
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x3F, 0x00, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x1F, 0x00, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "S":
        # class 3 (string): padding 1 = null padded (what h5py writes for numpy 'S'), ASCII
        return struct.pack("<BBBBI", 0x13, 0x01, 0x00, 0x00, max(1, dt.itemsize))
    raise TypeError("hdf5_lite: unsupported dtype %r" % (dt,))


def _dataspace_message(shape: Tuple[int, ...]) -> bytes:
    # version 1: version, rank, flags (no maximum sizes), 5 reserved bytes, then the dimension sizes
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


# ---------------------------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------------------------
class _Group:
    def __init__(self):
        self.children: Dict[str, Union["_Group", tuple]] = {}


class Writer:
    """Collects groups / datasets in memory (``w["a/b/c"] = array``) and lays the file out on ``close()``."""

    def __init__(self, path: str):
        self.path = path
        self.root = _Group()
        self._closed = False

    def __setitem__(self, name: str, value):
        parts = [p for p in name.split("/") if p]
        if not parts:
            raise ValueError("empty dataset name")
        g = self.root
        for p in parts[:-1]:
            nxt = g.children.setdefault(p, _Group())
            if not isinstance(nxt, _Group):
                raise ValueError("%r is a dataset, not a group" % p)
            g = nxt
        if parts[-1] in g.children:
            raise ValueError("%r exists already" % name)
        if isinstance(value, VlenStr):
            g.children[parts[-1]] = ("vlen", value)
        elif isinstance(value, (str, bytes)):
            g.children[parts[-1]] = ("vlen", VlenStr([value], shape=()))       # h5py stores a Python str as a scalar vlen string
        else:
            a = np.asarray(value)
            if a.ndim:                                                        # (ascontiguousarray would turn a scalar into [1])
                a = np.ascontiguousarray(a)
            if a.dtype.kind == "U":
                a = a.astype("S")
            if a.dtype.byteorder == ">":
                a = a.astype(a.dtype.newbyteorder("<"))
            _datatype_message(a.dtype)                                        # raises on unsupported types
            g.children[parts[-1]] = ("array", a)

    def require_group(self, name: str):
        g = self.root
        for p in [p for p in name.split("/") if p]:
            g = g.children.setdefault(p, _Group())

    # ---- layout -------------------------------------------------------------------------------------------------------
    def close(self):
        if self._closed:
            return
        self._closed = True
        # structures are streamed to the file as they are laid out (a whole-genome image file is gigabytes); the superblock
        # (56 bytes + the 40-byte root entry) and the B-tree sibling pointers are patched in afterwards
        with open(self.path, "wb") as f:
            self._f, self._pos = f, 96
            f.write(bytes(96))
            root_hdr, root_btree, root_heap = self._write_group(self.root)
            eof = self._pos
            sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
            sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
            sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
            assert len(sb) == 96
            f.seek(0)
            f.write(sb)
        self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _alloc(self, data: bytes) -> int:
        pad = (-self._pos) % 8
        if pad:
            self._f.write(bytes(pad))
            self._pos += pad
        at = self._pos
        self._f.write(data)
        self._pos += len(data)
        return at

    def _patch(self, at: int, data: bytes):
        self._f.seek(at)
        self._f.write(data)
        self._f.seek(self._pos)

    @staticmethod
    def _object_header(messages: List[Tuple[int, bytes, int]]) -> bytes:
        body = b""
        for mtype, data, flags in messages:
            data = data + b"\0" * (_pad8(len(data)) - len(data))
            body += struct.pack("<HHB3x", mtype, len(data), flags) + data
        # version 1, reserved, number of messages, reference count 1, header data size; 4 bytes of padding to 8
        return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body

    def _write_dataset(self, kind, value) -> int:
        if kind == "vlen":
            shape, raw = value.shape, self._write_vlen(value.values)
            dtype_msg = _datatype_message("vlen_str")
        else:
            shape, raw, dtype_msg = value.shape, (memoryview(value).cast("B") if value.ndim else value.tobytes()) if value.size else b"", _datatype_message(value.dtype)
        addr = self._alloc(raw) if raw else UNDEF
        msgs = [(MSG_DATASPACE, _dataspace_message(shape), 0),
                (MSG_DATATYPE, dtype_msg, 1),                                   # flag bit 0: constant message
                # fill value, version 1, byte for byte what the library wrote into the reader's reference file: allocate late (2),
                # write the fill value if one is set (2), "defined" with size 0 = the default fill value
                (MSG_FILL, struct.pack("<BBBBI", 1, 2, 2, 1, 0), 1),
                (MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, addr, len(raw)), 1)]    # v3, contiguous: address, size
        return self._alloc(self._object_header(msgs))

    def _write_vlen(self, values: List[bytes]) -> bytes:
        """the strings go to global heap collections; the dataset holds (length, collection address, object index) each"""
        out = bytearray()
        i = 0
        while i < len(values):
            # fill one collection: 16 bytes of header, objects of 16 + padded size, room for the free-space object 0
            # (an empty string is a null heap id -- length 0, address 0, index 0 -- like the library's "set null")
            objs, used, slots = [], 16, []
            while i < len(values) and (not objs or used + 16 + _pad8(len(values[i])) + 16 <= GCOL_MIN):
                if len(values[i]):
                    objs.append(values[i]); used += 16 + _pad8(len(values[i])); slots.append(len(objs))
                else:
                    slots.append(0)
                i += 1
            if not objs:
                out += b"".join(struct.pack("<IQI", 0, 0, 0) for _ in slots)
                continue
            size = max(GCOL_MIN, _pad8(used + 16))
            col = bytearray(b"GCOL" + struct.pack("<B3xQ", 1, size))
            for k, v in enumerate(objs):
                col += struct.pack("<HH4xQ", k + 1, 0, len(v)) + v + b"\0" * (_pad8(len(v)) - len(v))
            free = size - len(col)
            col += struct.pack("<HH4xQ", 0, 0, free) + b"\0" * (free - 16)      # object 0: the free space (size includes its header)
            addr = self._alloc(bytes(col))
            for k in slots:
                out += struct.pack("<IQI", len(objs[k - 1]), addr, k) if k else struct.pack("<IQI", 0, 0, 0)
        return bytes(out)

    def _write_group(self, g: _Group) -> Tuple[int, int, int]:
        """-> (object header address, B-tree address, local heap address)"""
        names = sorted(g.children, key=lambda s: s.encode("utf-8"))             # symbol tables are ordered by strcmp
        entries = []
        for n in names:
            c = g.children[n]
            if isinstance(c, _Group):
                hdr, bt, hp = self._write_group(c)
                entries.append((n, hdr, 1, struct.pack("<QQ", bt, hp)))          # cache type 1: B-tree + heap addresses cached
            else:
                entries.append((n, self._write_dataset(*c), 0, b"\0" * 16))
        # local heap: offset 0 holds the empty string (the B-tree's first key), then the names, 8-byte aligned
        heap = bytearray(8)
        off = {}
        for n in names:
            off[n] = len(heap)
            b = n.encode("utf-8") + b"\0"
            heap += b + b"\0" * (_pad8(len(b)) - len(b))
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 32) + b"\0" * 16                          # one free block: next = 1 (none), size 32
        heap_data = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        # symbol-table nodes of up to 2K entries each (allocated at full size, as the library does), then the B-tree
        per = 2 * GROUP_LEAF_K
        leaves = []
        for i in range(0, len(entries), per):
            chunk = entries[i:i + per]
            node = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)))
            for n, hdr, cache, scratch in chunk:
                node += struct.pack("<QQII", off[n], hdr, cache, 0) + scratch
            node += b"\0" * (8 + 40 * per - len(node))
            leaves.append((self._alloc(bytes(node)), off[chunk[-1][0]] if chunk else 0))
        bt = self._btree_level(leaves, 0)
        hdr = self._alloc(self._object_header([(MSG_SYMBOL_TABLE, struct.pack("<QQ", bt, heap_addr), 0)]))
        return hdr, bt, heap_addr

    def _btree_level(self, children: List[Tuple[int, int]], level: int) -> int:
        """children: (address, heap offset of the largest name below it). Nodes of up to 2K children; returns the root."""
        per = 2 * GROUP_INTERNAL_K
        nodes = []
        if not children:                                                         # an empty group: a leaf-level node without entries
            return self._alloc(b"TREE" + struct.pack("<BBH", 0, 0, 0) + struct.pack("<QQ", UNDEF, UNDEF) + b"\0" * (8 + 16 * per))
        groups = [children[i:i + per] for i in range(0, len(children), per)]
        addrs = []
        for gi, grp in enumerate(groups):
            body = bytearray()
            first_key = 0 if gi == 0 else groups[gi - 1][-1][1]
            body += struct.pack("<Q", first_key)
            for addr, key in grp:
                body += struct.pack("<QQ", addr, key)
            body += b"\0" * (8 + 16 * per - len(body))
            node = bytearray(b"TREE" + struct.pack("<BBH", 0, level, len(grp)) + struct.pack("<QQ", UNDEF, UNDEF)) + body
            addrs.append(self._alloc(bytes(node)))
            nodes.append((addrs[-1], grp[-1][1]))
        for gi, a in enumerate(addrs):                                           # sibling pointers
            left = addrs[gi - 1] if gi > 0 else UNDEF
            right = addrs[gi + 1] if gi + 1 < len(addrs) else UNDEF
            if left != UNDEF or right != UNDEF:
                self._patch(a + 8, struct.pack("<QQ", left, right))
        return nodes[0][0] if len(nodes) == 1 else self._btree_level(nodes, level + 1)


# ---------------------------------------------------------------------------------------------------------------------
# reader (strict about the structures above: signatures, versions, sizes; skips messages it does not need)
# ---------------------------------------------------------------------------------------------------------------------
class FormatError(ValueError):
    pass


class Reader:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.d = f.read()
        self.base = -1
        for off in [0] + [512 << k for k in range(16)]:                          # the superblock sits at 0, 512, 1024, ...
            if self.d[off:off + 8] == SIG:
                self.base = off
                break
        if self.base < 0:
            raise FormatError("no HDF5 signature")
        sb = self.d[self.base:]
        ver, fsv, rgv, _, shv, so, sl, _, self.leaf_k, self.int_k, self.flags = struct.unpack_from("<BBBBBBBBHHI", sb, 8)
        if ver != 0 or so != 8 or sl != 8:
            raise FormatError("superblock version %d / offsets %d / lengths %d not handled" % (ver, so, sl))
        base_addr, free_addr, self.eof, drv = struct.unpack_from("<QQQQ", sb, 24)
        if base_addr not in (0, self.base):
            raise FormatError("base address %d" % base_addr)
        if self.base + self.eof - base_addr > len(self.d) + 0 and self.eof - base_addr + self.base != len(self.d):
            raise FormatError("end-of-file address %d beyond the file (%d bytes)" % (self.eof, len(self.d)))
        name_off, hdr, cache, _ = struct.unpack_from("<QQII", sb, 56)
        self.root_header = hdr
        self.root = self._read_group(hdr)

    def _at(self, addr: int) -> int:
        if addr == UNDEF or self.base + addr > len(self.d):
            raise FormatError("address %d outside the file" % addr)
        return self.base + addr

    # ---- object headers --------------------------------------------------------------------------------------------------
    def _messages(self, addr: int):
        p = self._at(addr)
        ver, _, n_msgs, refs, size = struct.unpack_from("<BBHII", self.d, p)
        if ver != 1:
            raise FormatError("object header version %d at %d" % (ver, addr))
        blocks = [(p + 16, size)]
        out = []
        n_blocks = 0
        while blocks and len(out) < n_msgs:
            q, left = blocks.pop(0)
            n_blocks += 1
            if n_blocks > 64 or q + left > len(self.d):
                raise FormatError("object header at %d: continuation blocks" % addr)
            while left >= 8 and len(out) < n_msgs:
                mtype, msize, flags = struct.unpack_from("<HHB", self.d, q)
                data = self.d[q + 8:q + 8 + msize]
                if len(data) != msize or msize % 8:
                    raise FormatError("message of %d bytes at %d" % (msize, q))
                if mtype == MSG_CONTINUATION:
                    c_off, c_len = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self._at(c_off), c_len))
                out.append((mtype, data, flags))
                q += 8 + msize; left -= 8 + msize
        if len(out) != n_msgs:
            raise FormatError("object header at %d announces %d messages, %d found" % (addr, n_msgs, len(out)))
        return out

    # ---- groups ----------------------------------------------------------------------------------------------------------
    def _heap(self, addr: int):
        p = self._at(addr)
        if self.d[p:p + 4] != b"HEAP" or self.d[p + 4] != 0:
            raise FormatError("local heap at %d" % addr)
        size, free, data = struct.unpack_from("<QQQ", self.d, p + 8)
        q = self._at(data)
        seg = self.d[q:q + size]
        if len(seg) != size:
            raise FormatError("local heap data segment")
        f, steps = free, 0
        while f != 1:                                                            # walk the free list (1 = end of list)
            steps += 1
            if f + 16 > size or steps > size // 16 + 1:                          # (a cyclic list never ends)
                raise FormatError("local heap free block at %d" % f)
            f, fsz = struct.unpack_from("<QQ", seg, f)
            if fsz < 16:
                raise FormatError("local heap free block of %d bytes" % fsz)
        return seg

    def _name(self, heap: bytes, off: int) -> str:
        end = heap.index(b"\0", off)
        return heap[off:end].decode("utf-8")

    def _btree_leaves(self, addr: int, heap: bytes, want_level=None):
        p = self._at(addr)
        if self.d[p:p + 4] != b"TREE":
            raise FormatError("B-tree node at %d" % addr)
        ntype, level, used = struct.unpack_from("<BBH", self.d, p + 4)
        if ntype != 0 or used > 2 * self.int_k:
            raise FormatError("group B-tree node type %d with %d entries" % (ntype, used))
        if (want_level is not None and level != want_level) or level > 16:       # levels count down to the leaves: no cycles
            raise FormatError("B-tree node at %d has level %d" % (addr, level))
        keys = [struct.unpack_from("<Q", self.d, p + 24 + 16 * i)[0] for i in range(used + 1)]
        kids = [struct.unpack_from("<Q", self.d, p + 32 + 16 * i)[0] for i in range(used)]
        names = [self._name(heap, k) for k in keys]
        if any(names[i].encode() > names[i + 1].encode() for i in range(used)):
            raise FormatError("B-tree keys out of order at %d" % addr)
        out = []
        for i, c in enumerate(kids):
            out += [(c, names[i], names[i + 1])] if level == 0 else self._btree_leaves(c, heap, level - 1)
        return out

    def _read_group(self, hdr_addr: int):
        msgs = self._messages(hdr_addr)
        st = [m for m in msgs if m[0] == MSG_SYMBOL_TABLE]
        if not st:
            raise FormatError("object at %d is no old-style group" % hdr_addr)
        bt, hp = struct.unpack_from("<QQ", st[0][1], 0)
        heap = self._heap(hp)
        members = {}
        last = ""
        for leaf, lo, hi in self._btree_leaves(bt, heap):
            p = self._at(leaf)
            if self.d[p:p + 4] != b"SNOD" or self.d[p + 4] != 1:
                raise FormatError("symbol table node at %d" % leaf)
            n = struct.unpack_from("<H", self.d, p + 6)[0]
            if n > 2 * self.leaf_k:
                raise FormatError("symbol table node with %d entries" % n)
            for i in range(n):
                name_off, hdr, cache, _ = struct.unpack_from("<QQII", self.d, p + 8 + 40 * i)
                name = self._name(heap, name_off)
                if not (lo.encode() < name.encode() <= hi.encode()) and not (lo == "" and name.encode() <= hi.encode()):
                    raise FormatError("entry %r outside its B-tree key range (%r, %r]" % (name, lo, hi))
                if name.encode() <= last.encode() and last:
                    raise FormatError("symbol table entries out of order at %r" % name)
                last = name
                members[name] = hdr
        return members

    # ---- public ----------------------------------------------------------------------------------------------------------
    def keys(self, path: str = "/") -> List[str]:
        return sorted(self._resolve_group(path))

    def _resolve_group(self, path: str):
        g = self.root
        for p in [p for p in path.split("/") if p][:64]:
            if p not in g:
                raise KeyError(path)
            g = self._read_group(g[p])
        return g

    def is_group(self, path: str) -> bool:
        parts = [p for p in path.split("/") if p]
        g = self._resolve_group("/".join(parts[:-1]))
        return any(m[0] == MSG_SYMBOL_TABLE for m in self._messages(g[parts[-1]]))

    def describe(self, path: str) -> dict:
        """shape, datatype class / size / bit field and layout of a dataset (the decoded header messages)"""
        parts = [p for p in path.split("/") if p]
        g = self._resolve_group("/".join(parts[:-1]))
        if parts[-1] not in g:
            raise KeyError(path)
        info = {"messages": []}
        for mtype, data, flags in self._messages(g[parts[-1]]):
            info["messages"].append(mtype)
            if mtype == MSG_DATASPACE:
                ver, rank, fl = struct.unpack_from("<BBB", data, 0)
                off = 8 if ver == 1 else 4
                if ver not in (1, 2):
                    raise FormatError("dataspace version %d" % ver)
                info["shape"] = tuple(struct.unpack_from("<Q", data, off + 8 * i)[0] for i in range(rank))
            elif mtype == MSG_DATATYPE:
                cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", data, 0)
                info.update(type_class=cv & 15, type_version=cv >> 4, type_bits=(b0, b1, b2), type_size=size, type_props=bytes(data[8:]))
            elif mtype == MSG_LAYOUT:
                ver, cls = struct.unpack_from("<BB", data, 0)
                if ver in (1, 2):                   # the layout message of HDF5 1.6 and older (the reader's reference file has it)
                    rank, cls = data[1], data[2]
                    info["layout_class"] = cls
                    if cls == 1:
                        info["address"] = struct.unpack_from("<Q", data, 8)[0]
                        info["size"] = None         # product of the dimensions times the element size, filled in by the caller
                    elif cls == 0:
                        n = struct.unpack_from("<I", data, 8 + 4 * rank)[0]
                        info["compact"] = bytes(data[12 + 4 * rank:12 + 4 * rank + n])
                    continue
                if ver != 3:
                    raise FormatError("layout version %d" % ver)
                info["layout_class"] = cls
                if cls == 1:
                    info["address"], info["size"] = struct.unpack_from("<QQ", data, 2)
                elif cls == 0:
                    n = struct.unpack_from("<H", data, 2)[0]
                    info["compact"] = bytes(data[4:4 + n])
            elif mtype == MSG_FILL:
                info["fill"] = tuple(data[:4])
        return info

    def __getitem__(self, path: str):
        i = self.describe(path)
        for need in ("shape", "type_class", "layout_class"):
            if need not in i:
                raise FormatError("%s: no %s message" % (path, need.split("_")[0]))
        shape = i["shape"]
        n = 1
        for d in shape:
            n *= int(d)
        if n > len(self.d):                                       # more elements than the file has bytes
            raise FormatError("%s: %d elements in a file of %d bytes" % (path, n, len(self.d)))
        if i["layout_class"] == 1:
            if i["size"] is None:
                i["size"] = n * i["type_size"]
            raw = b"" if i["address"] == UNDEF else self.d[self._at(i["address"]):self._at(i["address"]) + i["size"]]
        elif i["layout_class"] == 0:
            raw = i["compact"]
        else:
            raise FormatError("chunked datasets are not handled")
        cls, size, bits = i["type_class"], i["type_size"], i["type_bits"]
        if (cls == 0 and size not in (1, 2, 4, 8)) or (cls == 1 and size not in (4, 8)) or (cls == 3 and not 0 < size <= (1 << 24)):
            raise FormatError("%s: datatype class %d of %d bytes" % (path, cls, size))
        if cls == 0:
            dt = np.dtype("%s%d" % ("i" if bits[0] & 8 else "u", size)).newbyteorder(">" if bits[0] & 1 else "<")
        elif cls == 1:
            dt = np.dtype("f%d" % size).newbyteorder(">" if bits[0] & 1 else "<")
        elif cls == 3:
            dt = np.dtype("S%d" % size)
        elif cls == 9 and (bits[0] & 15) == 1:
            out = []
            for k in range(n):
                ln, col, idx = struct.unpack_from("<IQI", raw, 16 * k)
                out.append(self._global_heap_object(col, idx)[:ln].decode("utf-8") if ln else "")
            return out[0] if not shape else np.array(out, dtype=object).reshape(shape)
        else:
            raise FormatError("datatype class %d not handled" % cls)
        if len(raw) != n * dt.itemsize:
            raise FormatError("%s: %d bytes stored, %d expected" % (path, len(raw), n * dt.itemsize))
        return np.frombuffer(raw, dt).reshape(shape)

    def _global_heap_object(self, col_addr: int, index: int) -> bytes:
        p = self._at(col_addr)
        if self.d[p:p + 4] != b"GCOL" or self.d[p + 4] != 1:
            raise FormatError("global heap collection at %d" % col_addr)
        size = struct.unpack_from("<Q", self.d, p + 8)[0]
        q, end = p + 16, p + size
        while q + 16 <= end:
            idx, refc, osz = struct.unpack_from("<HH4xQ", self.d, q)
            if idx == 0:
                if q + osz != end:
                    raise FormatError("global heap free space does not end the collection")
                break
            if idx == index:
                return self.d[q + 16:q + 16 + osz]
            q += 16 + _pad8(osz)
        raise FormatError("global heap object %d not found" % index)
