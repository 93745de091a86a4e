"""A minimal HDF5 writer / reader (pure Python + NumPy) for the saved-weights files of neural_network.py:188-196,
220-221: `wandb_anime_nn.h5` (model.save) and `wandb_main_weights.h5` (ModelCheckpoint).

h5py / libhdf5 are not part of this image, and the north-star wants the saved-weights layout kept intact, so the
subset of the HDF5 1.8 file format that h5py itself emits with its default `libver='earliest'` is written here
directly (HDF5 File Format Specification, version 1.1/2.0):

  superblock version 0 . old-style groups (object header v1 + symbol-table message -> v1 B-tree of group nodes +
  local heap) . datasets with contiguous layout (layout message v3), little-endian fixed-point / IEEE float /
  fixed-length string datatypes (datatype message v1), simple or scalar dataspaces (v1), fill-value message v2 .
  attributes as v1 attribute messages in the object header.

The reader handles the same subset plus object-header continuation blocks and compact layout, which is what a file
written by Keras through h5py contains for these small models; chunked or compressed datasets and variable-length
strings are reported as unsupported rather than guessed at.
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16          # group B-tree parameters stored in the superblock (the library defaults)
HEAP_FREE_NULL = 1                   # "end of free list" marker of a local heap on disk

MSG_DATASPACE, MSG_DATATYPE, MSG_FILL, MSG_LAYOUT, MSG_ATTR, MSG_CONT, MSG_SYMTAB = 0x1, 0x3, 0x5, 0x8, 0xC, 0x10, 0x11


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------ tree description
class Group:
    def __init__(self, attrs=None):
        self.children, self.attrs = {}, dict(attrs or {})

    def group(self, path):
        """The (created on demand) group at `path` below this one."""
        g = self
        for part in [p for p in path.split("/") if p]:
            g = g.children.setdefault(part, Group())
            if not isinstance(g, Group):
                raise ValueError("%r is a dataset" % part)
        return g

    def dataset(self, path, array, attrs=None):
        parts = [p for p in path.split("/") if p]
        g = self.group("/".join(parts[:-1]))
        g.children[parts[-1]] = Dataset(array, attrs)


class Dataset:
    def __init__(self, array, attrs=None):
        self.array, self.attrs = np.asarray(array), dict(attrs or {})


# ------------------------------------------------------------------------------------------ encoding
def _as_array(v):
    """Attribute / dataset value -> ndarray with a dtype this writer encodes."""
    if isinstance(v, str):
        v = v.encode("utf-8")
    if isinstance(v, (bytes, np.bytes_)):
        return np.array(bytes(v) or b"\x00", dtype="S%d" % max(1, len(v)))
    if isinstance(v, (list, tuple)) and v and all(isinstance(x, (bytes, str, np.bytes_)) for x in v):
        bs = [x.encode("utf-8") if isinstance(x, str) else bytes(x) for x in v]
        return np.array(bs, dtype="S%d" % max(1, max(len(b) for b in bs)))
    a = np.asarray(v)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    if a.dtype == np.float16 or a.dtype.kind == "b":
        raise TypeError("dtype %s is not supported by minih5" % a.dtype)
    return a


def _datatype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, man_bits, bias = (8, 23, 127) if dt.itemsize == 4 else (11, 52, 1023)
        head = struct.pack("<BBBBI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)       # null-padded, ASCII: what h5py writes for 'S'
    raise TypeError("dtype %s is not supported by minih5" % dt)


def _dataspace_msg(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _raw(a):
    a = np.ascontiguousarray(a)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return a.tobytes()


def _attr_msg(name, value):
    a = _as_array(value)
    nm = name.encode("utf-8") + b"\x00"
    dt, ds = _datatype_msg(a.dtype), _dataspace_msg(a.shape)
    return struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + _raw(a)


def _object_header(messages):
    body = b""
    for mtype, data in messages:
        data = _pad8(data)
        if len(data) > 0xFFFF:
            raise ValueError("object-header message of %d bytes does not fit (attribute too large)" % len(data))
        body += struct.pack("<HHB3x", mtype, len(data), 0) + data
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _File:
    def __init__(self):
        self.buf = bytearray(96)            # superblock, filled in at the end

    def put(self, data):
        self.buf += b"\x00" * (-len(self.buf) % 8)
        at = len(self.buf)
        self.buf += data
        return at


def _write_dataset(f, ds):
    a = _as_array(ds.array)
    raw = _raw(a)
    addr = f.put(raw) if raw else UNDEF
    msgs = [(MSG_DATASPACE, _dataspace_msg(a.shape)), (MSG_DATATYPE, _datatype_msg(a.dtype)),
            (MSG_FILL, struct.pack("<BBBB", 2, 1, 0, 0)),              # v2: early allocation, no fill value defined
            (MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, addr, len(raw)))]  # v3, contiguous
    msgs += [(MSG_ATTR, _attr_msg(k, v)) for k, v in ds.attrs.items()]
    return f.put(_object_header(msgs))


def _write_group(f, g):
    """-> (object header address, B-tree address, heap address)"""
    names = sorted(g.children, key=lambda s: s.encode("utf-8"))
    addrs = {}
    for n in names:
        c = g.children[n]
        addrs[n] = _write_group(f, c)[0] if isinstance(c, Group) else _write_dataset(f, c)
    # local heap: "" at offset 0, then the names, then one free block
    heap = bytearray(8)
    off = {}
    for n in names:
        off[n] = len(heap)
        heap += _pad8(n.encode("utf-8") + b"\x00")
    free_at = len(heap)
    heap += struct.pack("<QQ", HEAP_FREE_NULL, 32) + b"\x00" * 16
    heap_addr = f.put(b"")
    heap_addr = f.put(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), free_at, heap_addr + 32) + bytes(heap))
    # symbol-table nodes of at most 2*LEAF_K entries, in name order
    chunks = [names[i:i + 2 * LEAF_K] for i in range(0, len(names), 2 * LEAF_K)]     # none for an empty group
    if len(chunks) > 2 * INTERNAL_K:
        raise ValueError("group with %d members needs a two-level B-tree (not implemented)" % len(names))
    snods = []
    for ch in chunks:
        ent = b"".join(struct.pack("<QQII16x", off[n], addrs[n], 0, 0) for n in ch)
        ent += b"\x00" * (40 * (2 * LEAF_K - len(ch)))
        snods.append(f.put(struct.pack("<4sBBH", b"SNOD", 1, 0, len(ch)) + ent))
    keys = [0] + [off[ch[-1]] if ch else 0 for ch in chunks]
    node = struct.pack("<4sBBHQQ", b"TREE", 0, 0, len(chunks), UNDEF, UNDEF)
    for i, s in enumerate(snods):
        node += struct.pack("<QQ", keys[i], s)
    node += struct.pack("<Q", keys[len(snods)])
    node += b"\x00" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(node))
    btree = f.put(node)
    msgs = [(MSG_SYMTAB, struct.pack("<QQ", btree, heap_addr))]
    msgs += [(MSG_ATTR, _attr_msg(k, v)) for k, v in g.attrs.items()]
    return f.put(_object_header(msgs)), btree, heap_addr


def write(path, root):
    """Write the tree rooted at Group `root` as an HDF5 file."""
    f = _File()
    hdr, btree, heap = _write_group(f, root)
    f.buf += b"\x00" * (-len(f.buf) % 8)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(f.buf), UNDEF)
    sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    f.buf[:96] = sb
    with open(path, "wb") as fh:
        fh.write(f.buf)


# ------------------------------------------------------------------------------------------ reading
class Unsupported(Exception):
    pass


def _parse_datatype(b):
    cls, ver = b[0] & 0x0F, b[0] >> 4
    size = struct.unpack_from("<I", b, 4)[0]
    if ver not in (1, 2, 3):
        raise Unsupported("datatype message version %d" % ver)
    big = b[1] & 1
    if cls == 0:
        return np.dtype(("%s%s%d" % (">" if big else "<", "i" if b[1] & 0x08 else "u", size)))
    if cls == 1:
        return np.dtype("%sf%d" % (">" if big else "<", size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise Unsupported("datatype class %d (variable-length / compound / ...)" % cls)


def _parse_dataspace(b):
    ver, rank = b[0], b[1]
    if ver == 1:
        return tuple(struct.unpack_from("<%dQ" % rank, b, 8)) if rank else ()
    if ver == 2:
        if b[3] == 2:
            return None                   # null dataspace
        return tuple(struct.unpack_from("<%dQ" % rank, b, 4)) if rank else ()
    raise Unsupported("dataspace message version %d" % ver)


class _Reader:
    def __init__(self, data):
        self.d = data
        if data[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file")
        ver = data[8]
        if ver not in (0, 1):
            raise Unsupported("superblock version %d (written with libver='latest'?)" % ver)
        if data[13] != 8 or data[14] != 8:
            raise Unsupported("offset/length sizes other than 8 bytes")
        o = 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", data, o)[0]
        ent = o + 32
        self.root_header = struct.unpack_from("<Q", data, ent + 8)[0]

    def messages(self, addr):
        d = self.d
        ver = d[addr]
        if ver != 1:
            raise Unsupported("object header version %d" % ver)
        nmsg, _, size = struct.unpack_from("<HII", d, addr + 2)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", d, p)
                body = d[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == MSG_CONT:
                    off, ln = struct.unpack_from("<QQ", body)
                    blocks.append((off + self.base, ln))
                out.append((mtype, body))
        return out

    def attr(self, body):
        ver = body[0]
        if ver == 1:
            ns, ts, ss = struct.unpack_from("<HHH", body, 2)
            p = 8
            name = body[p:p + ns].split(b"\x00")[0].decode("utf-8")
            p += (ns + 7) // 8 * 8
            dt = body[p:p + ts]
            p += (ts + 7) // 8 * 8
            sp = body[p:p + ss]
            p += (ss + 7) // 8 * 8
        elif ver in (2, 3):
            ns, ts, ss = struct.unpack_from("<HHH", body, 2)
            p = 8 + (1 if ver == 3 else 0)
            name = body[p:p + ns].split(b"\x00")[0].decode("utf-8")
            p += ns
            dt = body[p:p + ts]
            p += ts
            sp = body[p:p + ss]
            p += ss
        else:
            raise Unsupported("attribute message version %d" % ver)
        try:
            dtype, shape = _parse_datatype(dt), _parse_dataspace(sp)
        except Unsupported:
            return name, None
        if shape is None:
            return name, None
        n = int(np.prod(shape)) if shape else 1
        return name, np.frombuffer(body, dtype, n, p).reshape(shape).copy()

    def node(self, addr):
        """-> ("group", {name: address}, attrs) or ("dataset", ndarray, attrs)"""
        msgs = self.messages(addr)
        attrs = {}
        for t, b in msgs:
            if t == MSG_ATTR:
                k, v = self.attr(b)
                attrs[k] = v
        sym = [b for t, b in msgs if t == MSG_SYMTAB]
        if sym:
            btree, heap = struct.unpack_from("<QQ", sym[0])
            return "group", self.group_entries(btree + self.base, heap + self.base), attrs
        dt = [b for t, b in msgs if t == MSG_DATATYPE]
        sp = [b for t, b in msgs if t == MSG_DATASPACE]
        lay = [b for t, b in msgs if t == MSG_LAYOUT]
        if not (dt and sp and lay):
            raise Unsupported("object at %d is neither an old-style group nor a dataset" % addr)
        dtype, shape = _parse_datatype(dt[0]), _parse_dataspace(sp[0])
        n = int(np.prod(shape)) if shape else 1
        lb = lay[0]
        if lb[0] != 3:
            raise Unsupported("data layout message version %d" % lb[0])
        if lb[1] == 1:
            a, size = struct.unpack_from("<QQ", lb, 2)
            arr = np.zeros(shape, dtype) if (a == UNDEF or n == 0) else \
                np.frombuffer(self.d, dtype, n, a + self.base).reshape(shape).copy()
        elif lb[1] == 0:
            size = struct.unpack_from("<H", lb, 2)[0]
            arr = np.frombuffer(lb, dtype, n, 4).reshape(shape).copy()
        else:
            raise Unsupported("chunked dataset layout")
        return "dataset", arr, attrs

    def group_entries(self, btree, heap):
        d = self.d
        if d[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        seg = struct.unpack_from("<Q", d, heap + 24)[0] + self.base
        out = {}

        def walk(a):
            if d[a:a + 4] == b"TREE":
                level, used = d[a + 5], struct.unpack_from("<H", d, a + 6)[0]
                for i in range(used):
                    child = struct.unpack_from("<Q", d, a + 24 + 8 + 16 * i)[0]
                    walk(child + self.base)
            elif d[a:a + 4] == b"SNOD":
                n = struct.unpack_from("<H", d, a + 6)[0]
                for i in range(n):
                    no, oh = struct.unpack_from("<QQ", d, a + 8 + 40 * i)
                    end = d.index(b"\x00", seg + no)
                    out[d[seg + no:end].decode("utf-8")] = oh + self.base
            else:
                raise ValueError("bad group node signature at %d" % a)
        walk(btree)
        return out


def read(path):
    """-> (datasets: {"/a/b": ndarray}, attrs: {"/a": {name: value}})  ("" is the root group)."""
    with open(path, "rb") as fh:
        r = _Reader(fh.read())
    datasets, attrs = {}, {}

    def visit(addr, path):
        kind, payload, a = r.node(addr)
        attrs[path] = a
        if kind == "group":
            for name, child in payload.items():
                visit(child, path + "/" + name)
        else:
            datasets[path] = payload
    visit(r.root_header + r.base, "")
    return datasets, attrs
