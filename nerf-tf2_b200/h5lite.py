"""
A minimal HDF5 reader/writer for Keras weight files.

The reference checkpoints its two sub-models with `Model.save_weights("*.h5")` (core/ops.py:142-143)
and restores them with `load_weights` (core/model.py:262-276). There is no HDF5 library in this image
(no h5py, no libhdf5), so this module implements the part of the HDF5 file format that such files use,
from the published format specification ("HDF5 File Format Specification Version 2.0/3.0"):

  superblock v0/v1 (optionally behind a user block), old-style groups (symbol-table message ->
  v1 B-tree "TREE" -> symbol nodes "SNOD" -> local heap "HEAP"), v1 object headers with continuation
  blocks, dataspace v1/v2, datatypes fixed-point / IEEE float / fixed-length string / variable-length
  string (global heap "GCOL"), data layout v3 compact / contiguous (chunked only without filters),
  attribute messages v1-v3.

That is what libhdf5 writes with its default ("earliest") format bounds, which is what h5py uses for
`save_weights`. New-style groups (fractal heaps, v2 B-trees), filters/compression and v2 object headers
("OHDR") are rejected with a clear error, never guessed at.

The READER is checked against a file written by the real libhdf5 (the MATLAB v7.3 sample shipped in
SciPy's test data) in tests/test_checkpoint_h5.py. The WRITER emits the same old-style structures; it is
verified by round trip through the reader and by structural checks only -- no libhdf5 exists here to
confirm that h5py accepts its files, and the module says so rather than claiming compatibility.
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ====================================================================== reader
class Dataset:
    def __init__(self, name, value, attrs):
        self.name, self.value, self.attrs = name, value, attrs


class Group:
    def __init__(self, name, attrs):
        self.name, self.attrs, self.children = name, attrs, {}

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            node = node.children[part]
        return node

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def keys(self):
        return list(self.children.keys())

    def visit_datasets(self, prefix=""):
        """Yields (path, ndarray) depth first, children in stored (name-sorted) order."""
        for k, v in self.children.items():
            p = f"{prefix}/{k}" if prefix else k
            if isinstance(v, Group):
                yield from v.visit_datasets(p)
            else:
                yield p, v.value


class _Reader:
    def __init__(self, buf):
        self.b = buf
        start = -1
        off = 0
        while off < len(buf):                       # the superblock sits at 0, 512, 1024, 2048, ...
            if buf[off:off + 8] == SIGNATURE:
                start = off
                break
            off = 512 if off == 0 else off * 2
        if start < 0:
            raise H5Error("not an HDF5 file (signature not found)")
        ver = buf[start + 8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} (new-style file) is not supported; only the default "
                          "'earliest' layout written by h5py/Keras is")
        self.O, self.L = buf[start + 13], buf[start + 14]
        if self.O != 8 or self.L != 8:
            raise H5Error(f"offset/length sizes {self.O}/{self.L} are not supported (expected 8/8)")
        p = start + 24 + (4 if ver == 1 else 0)
        self.base = self.u64(p)
        if self.base == UNDEF:
            self.base = 0
        if start and self.base == 0:
            self.base = start                     # some writers leave 0 with a user block: addresses are relative to it
        p += 32                                    # base, free-space, eof, driver-info
        self.root_entry = self._symbol_entry(p)

    # ---- primitives
    def u8(self, p):
        return self.b[p]

    def u16(self, p):
        return struct.unpack_from("<H", self.b, p)[0]

    def u32(self, p):
        return struct.unpack_from("<I", self.b, p)[0]

    def u64(self, p):
        return struct.unpack_from("<Q", self.b, p)[0]

    def addr(self, p):
        a = self.u64(p)
        return None if a == UNDEF else a + self.base

    def _symbol_entry(self, p):
        e = {"name_off": self.u64(p), "header": self.addr(p + 8), "cache": self.u32(p + 16)}
        if e["cache"] == 1:
            e["btree"], e["heap"] = self.addr(p + 24), self.addr(p + 32)
        return e

    # ---- object headers
    def messages(self, header_addr):
        """[(type, flags, payload offset, size)] of a v1 object header, continuation blocks followed."""
        b = self.b
        if b[header_addr:header_addr + 4] == b"OHDR":
            raise H5Error("version-2 object headers (libver='latest' files) are not supported")
        if b[header_addr] != 1:
            raise H5Error(f"object header version {b[header_addr]} at {header_addr:#x} not supported")
        n_msgs = self.u16(header_addr + 2)
        size = self.u32(header_addr + 8)
        blocks = [(header_addr + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            p, remaining = blocks.pop(0)
            end = p + remaining
            while p + 8 <= end and len(out) < n_msgs:
                mtype, msize, flags = self.u16(p), self.u16(p + 2), self.u8(p + 4)
                body = p + 8
                if mtype == 0x0010:                 # continuation
                    blocks.append((self.addr(body), self.u64(body + 8)))
                out.append((mtype, flags, body, msize))
                p = body + msize
        return out

    # ---- datatypes / dataspaces
    def datatype(self, p):
        """-> (descriptor, encoded size). descriptor: ('num', dtype) | ('str', n) | ('vlen_str',) | ('vlen', base)"""
        cv = self.u8(p)
        cls, ver = cv & 0x0F, cv >> 4
        bits0 = self.u8(p + 1)
        size = self.u32(p + 4)
        if cls == 0:                                # fixed point: byte order bit0, signed bit3
            order = ">" if bits0 & 1 else "<"
            kind = "i" if bits0 & 8 else "u"
            return ("num", np.dtype(f"{order}{kind}{size}")), 8 + 4
        if cls == 1:                                # IEEE float
            order = ">" if bits0 & 1 else "<"
            if size not in (2, 4, 8):
                raise H5Error(f"float of {size} bytes not supported")
            return ("num", np.dtype(f"{order}f{size}")), 8 + 12
        if cls == 3:                                # fixed-length string
            return ("str", size), 8
        if cls == 9:                                # variable length
            base, blen = self.datatype(p + 8)
            if (bits0 & 0x0F) == 1:
                return ("vlen_str",), 8 + blen
            return ("vlen", base), 8 + blen
        raise H5Error(f"datatype class {cls} (version {ver}) not supported")

    def dataspace(self, p):
        ver, rank, flags = self.u8(p), self.u8(p + 1), self.u8(p + 2)
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if self.u8(p + 3) == 2:                 # null dataspace
                return None
            q = p + 4
        else:
            raise H5Error(f"dataspace version {ver} not supported")
        return tuple(self.u64(q + 8 * i) for i in range(rank))

    def _global_heap_object(self, coll_addr, index):
        b = self.b
        if b[coll_addr:coll_addr + 4] != b"GCOL":
            raise H5Error("bad global heap collection signature")
        size = self.u64(coll_addr + 8)
        p, end = coll_addr + 16, coll_addr + size
        while p + 16 <= end:
            idx, osize = self.u16(p), self.u64(p + 8)
            if idx == 0:
                break
            if idx == index:
                return bytes(b[p + 16:p + 16 + osize])
            p += 16 + _pad8(osize)
        raise H5Error(f"global heap object {index} not found")

    def decode(self, desc, shape, raw):
        n = 1 if shape is None else int(np.prod(shape, dtype=np.int64))
        if shape is None:
            return None
        if desc[0] == "num":
            arr = np.frombuffer(raw, dtype=desc[1], count=n).reshape(shape)
            return arr.astype(desc[1].newbyteorder("=")) if shape else arr.astype(desc[1].newbyteorder("="))[()]
        if desc[0] == "str":
            arr = np.frombuffer(raw, dtype=f"S{desc[1]}", count=n).reshape(shape)
            return arr.copy() if shape else arr[()]
        if desc[0] == "vlen_str":
            vals = []
            for i in range(n):
                q = 16 * i
                ln = struct.unpack_from("<I", raw, q)[0]
                coll = struct.unpack_from("<Q", raw, q + 4)[0]
                idx = struct.unpack_from("<I", raw, q + 12)[0]
                vals.append(self._global_heap_object(coll + self.base, idx)[:ln] if ln else b"")
            arr = np.array(vals, dtype=object).reshape(shape)
            return arr if shape else arr[()]
        raise H5Error(f"cannot decode datatype {desc}")

    def _itemsize(self, desc):
        return {"num": lambda: desc[1].itemsize, "str": lambda: desc[1], "vlen_str": lambda: 16, "vlen": lambda: 16}[desc[0]]()

    # ---- attributes
    def attribute(self, p):
        ver = self.u8(p)
        name_sz, dt_sz, ds_sz = self.u16(p + 2), self.u16(p + 4), self.u16(p + 6)
        if ver == 1:
            q = p + 8
            name = bytes(self.b[q:q + name_sz]).split(b"\0")[0].decode()
            q += _pad8(name_sz)
            desc, _ = self.datatype(q)
            q += _pad8(dt_sz)
            shape = self.dataspace(q)
            q += _pad8(ds_sz)
        elif ver in (2, 3):
            q = p + 8 + (1 if ver == 3 else 0)
            name = bytes(self.b[q:q + name_sz]).split(b"\0")[0].decode()
            q += name_sz
            desc, _ = self.datatype(q)
            q += dt_sz
            shape = self.dataspace(q)
            q += ds_sz
        else:
            raise H5Error(f"attribute message version {ver} not supported")
        n = 0 if shape is None else int(np.prod(shape, dtype=np.int64))
        raw = bytes(self.b[q:q + n * self._itemsize(desc)])
        return name, self.decode(desc, shape, raw)

    # ---- groups
    def _heap_name(self, heap_addr, off):
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        data = self.addr(heap_addr + 24)
        end = self.b.index(b"\0", data + off)
        return bytes(self.b[data + off:end]).decode()

    def _btree_entries(self, node, heap):
        b = self.b
        if b[node:node + 4] == b"SNOD":
            n = self.u16(node + 6)
            for i in range(n):
                e = self._symbol_entry(node + 8 + 40 * i)
                yield self._heap_name(heap, e["name_off"]), e
            return
        if b[node:node + 4] != b"TREE":
            raise H5Error(f"bad group B-tree signature at {node:#x}")
        if self.u8(node + 4) != 0:
            raise H5Error("unexpected B-tree node type in a group")
        used = self.u16(node + 6)
        p = node + 24
        for i in range(used):
            child = self.addr(p + 8 + 16 * i)
            yield from self._btree_entries(child, heap)

    def read_object(self, name, header_addr):
        msgs = self.messages(header_addr)
        attrs = {}
        symtab = layout = dtype = None
        shape = ()
        for mtype, flags, p, size in msgs:
            if mtype == 0x000C:
                k, v = self.attribute(p)
                attrs[k] = v
            elif mtype == 0x0011:
                symtab = (self.addr(p), self.addr(p + 8))
            elif mtype == 0x0001:
                shape = self.dataspace(p)
            elif mtype == 0x0003:
                dtype, _ = self.datatype(p)
            elif mtype == 0x0008:
                layout = p
            elif mtype == 0x000B:
                raise H5Error(f"{name}: filtered (compressed) datasets are not supported")
            elif mtype in (0x0002, 0x0006):
                raise H5Error(f"{name}: new-style groups (link messages) are not supported")
        if symtab is not None:
            g = Group(name, attrs)
            for child_name, e in self._btree_entries(symtab[0], symtab[1]):
                g.children[child_name] = self.read_object(child_name, e["header"])
            return g
        if layout is None or dtype is None:
            raise H5Error(f"{name}: neither a group nor a dataset")
        return Dataset(name, self._dataset_value(name, layout, dtype, shape), attrs)

    def _dataset_value(self, name, p, desc, shape):
        ver = self.u8(p)
        n = 0 if shape is None else int(np.prod(shape, dtype=np.int64))
        nbytes = n * self._itemsize(desc)
        if ver in (1, 2):                           # pre-1.6.3 layout message: rank, class, 5 reserved, address, dims
            rank, cls = self.u8(p + 1), self.u8(p + 2)
            if cls == 1:
                a = self.addr(p + 8)
                return self.decode(desc, shape, b"\0" * nbytes if a is None else bytes(self.b[a:a + nbytes]))
            if cls == 0:
                q = p + 8 + 4 * rank
                return self.decode(desc, shape, bytes(self.b[q + 4:q + 4 + self.u32(q)]))
            raise H5Error(f"{name}: layout v{ver} class {cls} not supported")
        if ver != 3:
            raise H5Error(f"{name}: data layout message version {ver} not supported")
        cls = self.u8(p + 1)
        if cls == 0:                                # compact
            size = self.u16(p + 2)
            raw = bytes(self.b[p + 4:p + 4 + size])
        elif cls == 1:                              # contiguous
            a = self.addr(p + 2)
            raw = b"\0" * nbytes if a is None else bytes(self.b[a:a + nbytes])
        elif cls == 2:
            raw = self._chunked(name, p, desc, shape)
        else:
            raise H5Error(f"{name}: layout class {cls} not supported")
        return self.decode(desc, shape, raw)

    def _chunked(self, name, p, desc, shape):
        rank = self.u8(p + 2)                       # dataset rank + 1
        btree = self.addr(p + 3)
        cdims = [self.u32(p + 11 + 4 * i) for i in range(rank)]
        item = cdims[-1]
        cshape = tuple(cdims[:-1])
        out = np.zeros(shape, dtype=np.uint8).reshape(shape + (1,)).repeat(item, axis=-1)

        def walk(node):
            if self.b[node:node + 4] != b"TREE" or self.u8(node + 4) != 1:
                raise H5Error(f"{name}: bad chunk B-tree")
            level, used = self.u8(node + 5), self.u16(node + 6)
            key_sz = 8 + 8 * rank
            q = node + 24
            for i in range(used):
                k = q + i * (key_sz + 8)
                csize, mask = self.u32(k), self.u32(k + 4)
                offs = [self.u64(k + 8 + 8 * d) for d in range(rank - 1)]
                child = self.addr(k + key_sz)
                if level:
                    walk(child)
                    continue
                if mask:
                    raise H5Error(f"{name}: filtered chunks are not supported")
                chunk = np.frombuffer(self.b, dtype=np.uint8, count=csize, offset=child).reshape(cshape + (item,))
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        if btree is not None:
            walk(btree)
        return out.tobytes()


def read_file(path):
    """Parse a whole (small) HDF5 file into a tree of Group / Dataset objects with decoded attributes."""
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(buf)
    return r.read_object("/", r.root_entry["header"])


def load_keras_weights(path):
    """-> ordered [(weight name, ndarray)] of a Keras `save_weights` file, in `layer_names` x
    `weight_names` order (keras/saving/hdf5_format.py: save_weights_to_hdf5_group)."""
    root = read_file(path)
    if "model_weights" in root.children:            # a full-model `model.save(...)` file
        root = root["model_weights"]
    out = []
    for layer in root.attrs["layer_names"]:
        g = root[layer.decode()]
        for w in g.attrs.get("weight_names", []):
            out.append((w.decode(), np.asarray(g[w.decode()].value)))
    return out


# ====================================================================== writer (old-style structures)
class _Writer:
    """Lays objects out depth first into one bytearray. Groups: one SNOD per group (<= 2*K entries,
    K = 16 by default -> up to 32 children, plenty for a 13-layer model), names sorted as libhdf5 keeps them."""

    LEAF_K = 16

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, n, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        a = len(self.buf)
        self.buf.extend(b"\0" * n)
        return a

    def put(self, a, data):
        self.buf[a:a + len(data)] = data

    # ---- message bodies
    @staticmethod
    def dt_msg(arr):
        if arr.dtype.kind == "f":
            size = arr.dtype.itemsize
            exp_loc, exp_sz, man_sz, bias = {2: (10, 5, 10, 15), 4: (23, 8, 23, 127), 8: (52, 11, 52, 1023)}[size]
            bits = bytes([0x20, size * 8 - 1, 0x00])                # little endian, mantissa norm = implied msb, sign position
            props = struct.pack("<HHBBBBI", 0, size * 8, exp_loc, exp_sz, 0, man_sz, bias)
            return bytes([0x11]) + bits + struct.pack("<I", size) + props
        if arr.dtype.kind in "iu":
            size = arr.dtype.itemsize
            bits = bytes([0x08 if arr.dtype.kind == "i" else 0x00, 0, 0])
            return bytes([0x10]) + bits + struct.pack("<I", size) + struct.pack("<HH", 0, size * 8)
        if arr.dtype.kind == "S":
            return bytes([0x13, 0x01, 0, 0]) + struct.pack("<I", arr.dtype.itemsize)    # null-padded (NumPy 'S'), ASCII
        raise H5Error(f"cannot write dtype {arr.dtype}")

    @staticmethod
    def ds_msg(shape):
        body = bytes([1, len(shape), 0, 0, 0, 0, 0, 0])
        return body + b"".join(struct.pack("<Q", int(d)) for d in shape)

    @classmethod
    def attr_msg(cls, name, arr):
        arr = np.require(arr, requirements="C")       # (ascontiguousarray would turn a scalar into shape (1,))
        nm = name.encode() + b"\0"
        dt, ds = cls.dt_msg(arr), cls.ds_msg(arr.shape)
        body = bytes([1, 0]) + struct.pack("<HHH", len(nm), len(dt), len(ds))
        for part in (nm, dt, ds):
            body += part + b"\0" * (_pad8(len(part)) - len(part))
        return body + arr.tobytes()

    def object_header(self, msgs):
        parts = b""
        for mtype, body in msgs:
            body = body + b"\0" * (_pad8(len(body)) - len(body))
            parts += struct.pack("<HHBBBB", mtype, len(body), 0, 0, 0, 0) + body
        a = self.alloc(16 + len(parts))
        self.put(a, struct.pack("<BBHII", 1, 0, len(msgs), 1, len(parts)) + b"\0\0\0\0" + parts)
        return a

    def dataset(self, arr, attrs):
        arr = np.require(arr, requirements="C")
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        data = self.alloc(max(arr.nbytes, 1))
        self.put(data, arr.tobytes())
        layout = bytes([3, 1]) + struct.pack("<QQ", data, arr.nbytes)
        fill = bytes([2, 2, 2, 0])                    # fill value v2: late allocation, write if set, undefined
        msgs = [(0x0001, self.ds_msg(arr.shape)), (0x0003, self.dt_msg(arr)), (0x0005, fill), (0x0008, layout)]
        msgs += [(0x000C, self.attr_msg(k, v)) for k, v in attrs.items()]
        return self.object_header(msgs)

    def group(self, children, attrs):
        """children: {name: ("group", children, attrs) | ("dataset", array, attrs)} -> (header, btree, heap)"""
        names = sorted(children)
        if len(names) > 2 * self.LEAF_K:
            raise H5Error("too many children for the single-node group writer")
        headers = {}
        for n in names:
            kind, payload, cattrs = children[n]
            headers[n] = self.group(payload, cattrs) if kind == "group" else (self.dataset(payload, cattrs), None, None)
        # local heap: offset 0 holds the empty string (first B-tree key)
        heap_data = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            enc = n.encode() + b"\0"
            heap_data += enc + b"\0" * (_pad8(len(enc)) - len(enc))
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)                      # one free block: next = 1 (none), size 16
        dseg = self.alloc(len(heap_data))
        self.put(dseg, heap_data)
        heap = self.alloc(32)
        self.put(heap, b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap_data), free_off, dseg))
        snod = self.alloc(8 + 40 * 2 * self.LEAF_K)
        body = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(names))
        for n in names:
            h, bt, hp = headers[n]
            if bt is None:
                body += struct.pack("<QQII", offs[n], h, 0, 0) + b"\0" * 16
            else:
                body += struct.pack("<QQII", offs[n], h, 1, 0) + struct.pack("<QQ", bt, hp)
        self.put(snod, body)
        btree = self.alloc(24 + (2 * self.LEAF_K + 1) * 8 + 2 * self.LEAF_K * 8)
        last = offs[names[-1]] if names else 0
        self.put(btree, b"TREE" + bytes([0, 0]) + struct.pack("<H", 1 if names else 0) +
                 struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, last))
        msgs = [(0x0011, struct.pack("<QQ", btree, heap))]
        msgs += [(0x000C, self.attr_msg(k, v)) for k, v in attrs.items()]
        return self.object_header(msgs), btree, heap


def write_file(path, children, attrs=None):
    """Write a tree {name: ("group", {...}, attrs) | ("dataset", ndarray, attrs)} as an old-style HDF5 file."""
    w = _Writer()
    w.alloc(96)                                                    # superblock v0 + root symbol-table entry
    header, btree, heap = w.group(children, attrs or {})
    eof = len(w.buf)
    sb = SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", _Writer.LEAF_K, 16, 0)   # group leaf K (symbol nodes hold 2K entries), internal K
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
    w.put(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(w.buf))


def save_keras_weights(path, layers, backend=b"tensorflow", keras_version=b"2.7.0"):
    """`layers` = ordered [(layer name, [(weight name, ndarray), ...])] -> a file with the structure of
    Keras `save_weights` (keras/saving/hdf5_format.py save_weights_to_hdf5_group): root attributes
    layer_names / backend / keras_version, one group per layer (a '/' in the name nests groups, as
    h5py's create_group does) carrying `weight_names`, datasets at <layer group>/<weight name>.
    Scalar string attributes are written as fixed-length strings (h5py would use variable-length ones)."""
    def fixed(strings):
        strings = [s.encode() if isinstance(s, str) else s for s in strings]
        return np.array(strings, dtype=f"S{max([len(s) for s in strings] + [1])}")

    def descend(children, parts):
        for part in parts:
            children = children.setdefault(part, ("group", {}, {}))[1]
        return children

    root = {}
    for lname, weights in layers:
        parts = lname.split("/")
        parent = descend(root, parts[:-1])
        node = parent.setdefault(parts[-1], ("group", {}, {}))
        node[2]["weight_names"] = fixed([w for w, _ in weights]) if weights else np.zeros((0,), dtype="S1")
        for wname, arr in weights:
            wparts = wname.split("/")
            descend(node[1], wparts[:-1])[wparts[-1]] = ("dataset", np.asarray(arr), {})
    attrs = {"layer_names": fixed([ln for ln, _ in layers]), "backend": fixed([backend]).reshape(()),
             "keras_version": fixed([keras_version]).reshape(())}
    write_file(path, root, attrs)
