"""Keras weight files (`model.save_weights(...h5)`, `load_weights(path, by_name=True)`: mpunet/models/
model_init.py:31,56, bin/train.py:303-317, bin/predict.py:205-233) without h5py.

A small reader / writer for the subset of the HDF5 file format those files use, written from the public
"HDF5 File Format Specification" (version-0/1 superblock, version-1 object headers, symbol-table groups = v1 B-tree
+ local heap + SNOD nodes, contiguous / compact / unfiltered chunked dataset layouts, version 1-3 attribute
messages, fixed-length string, integer and IEEE float datatypes).  That is what h5py writes with its default
`libver='earliest'` and what Keras 2.3 relies on:

    /                       attrs: layer_names (array of fixed-length byte strings), backend, keras_version
    /<layer>/               attrs: weight_names (e.g. b"encoder_L0_conv1/kernel:0")
    /<layer>/<layer>/kernel:0 ...   float32 datasets
    (full-model files written by model.save() keep the same tree under /model_weights)

STATUS: h5py / libhdf5 do not exist in the build environment, so the reader has only been exercised against files
produced by the writer below (round trips, tests/test_keras_h5.py) - it has NOT yet been validated against a file
written by h5py itself.  Not supported: compressed or otherwise filtered chunks, variable-length strings, new-style
(link-message / fractal-heap) groups, version-2 object headers; they raise NotImplementedError with the feature
named.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


def _pad8(n):
    return (n + 7) & ~7


# =====================================================================================================
# reader
# =====================================================================================================
class VLenStr(object):
    """Marks an attribute value the writer should store as a variable-length string (h5py's encoding of str/bytes)."""

    def __init__(self, value):
        self.value = value


class _Datatype(object):
    def __init__(self, cls, size, np_dtype=None, strpad=None):
        self.cls, self.size, self.np_dtype, self.strpad = cls, size, np_dtype, strpad


class _UnsupportedAttr(object):
    """Placeholder for an attribute whose datatype this subset reader cannot decode: opening the file must not
    fail because of it (Keras writes `backend` / `keras_version` next to the weights), only READING it does."""

    def __init__(self, name, why):
        self.name, self.why = name, why


class _Attrs(dict):
    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, _UnsupportedAttr):
            raise NotImplementedError("attribute %r: %s" % (v.name, v.why))
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default


class H5Node(object):
    def __init__(self, f, addr):
        self._f = f
        self._addr = addr
        self._msgs = f._read_object_header(addr)
        self.attrs = _Attrs()
        for mtype, data in self._msgs:
            if mtype == 0x000C:
                try:
                    name, value = f._parse_attribute(data)
                except NotImplementedError as e:
                    name = f._attribute_name(data)
                    value = _UnsupportedAttr(name, str(e))
                self.attrs[name] = value


class H5Dataset(H5Node):
    def __init__(self, f, addr):
        super().__init__(f, addr)
        m = dict((t, d) for t, d in self._msgs)
        if 0x0001 not in m or 0x0003 not in m or 0x0008 not in m:
            raise ValueError("object at %d is not a dataset" % addr)
        if 0x000B in m:
            raise NotImplementedError("HDF5 feature not supported: filtered (compressed) datasets")
        self.shape = f._parse_dataspace(m[0x0001])
        self._dt = f._parse_datatype(m[0x0003])
        self.dtype = self._dt.np_dtype
        self._layout = m[0x0008]

    def read(self):
        f, lay = self._f, self._layout
        n = int(np.prod(self.shape)) if len(self.shape) else 1
        nbytes = n * self._dt.size
        version, cls = lay[0], lay[1]
        if version != 3:
            raise NotImplementedError("HDF5 feature not supported: data layout message version %d" % version)
        if cls == 0:  # compact
            size = struct.unpack_from("<H", lay, 2)[0]
            raw = lay[4:4 + size]
        elif cls == 1:  # contiguous
            addr, size = struct.unpack_from("<QQ", lay, 2)
            raw = b"\0" * nbytes if addr == UNDEF else f.buf[addr:addr + nbytes]
        elif cls == 2:  # chunked, no filters
            rank1 = lay[2]
            btree = struct.unpack_from("<Q", lay, 3)[0]
            cdims = struct.unpack_from("<%dI" % rank1, lay, 11)
            out = np.zeros(self.shape, dtype=self.dtype)
            if btree != UNDEF:
                for offs, caddr, csize in f._iter_chunks(btree, rank1):
                    chunk = np.frombuffer(f.buf, dtype=self.dtype, count=int(np.prod(cdims[:-1])),
                                          offset=caddr).reshape(cdims[:-1])
                    sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims[:-1], self.shape))
                    sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
                    out[sl_out] = chunk[sl_in]
            return out
        else:
            raise NotImplementedError("HDF5 feature not supported: layout class %d" % cls)
        return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()


class H5Group(H5Node):
    def __init__(self, f, addr):
        super().__init__(f, addr)
        self._links = None

    def _load(self):
        if self._links is None:
            st = [d for t, d in self._msgs if t == 0x0011]
            if not st:
                if any(t in (0x0002, 0x0006) for t, _ in self._msgs):
                    raise NotImplementedError("HDF5 feature not supported: new-style groups (link messages)")
                raise ValueError("object at %d is not a group" % self._addr)
            btree, heap = struct.unpack_from("<QQ", st[0], 0)
            self._links = dict(self._f._iter_group(btree, heap))
        return self._links

    def keys(self):
        return sorted(self._load())

    def __contains__(self, name):
        return name in self._load()

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            links = node._load()
            if part not in links:
                raise KeyError(path)
            node = node._f._open(links[part])
        return node


class H5File(H5Group):
    """Read-only view of an HDF5 file held in memory."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        base = self.buf.find(SIGNATURE)
        if base != 0:
            raise ValueError("%s is not an HDF5 file (signature not found at offset 0)" % path)
        version = self.buf[8]
        if version not in (0, 1):
            raise NotImplementedError("HDF5 feature not supported: superblock version %d" % version)
        self.size_offsets, self.size_lengths = self.buf[13], self.buf[14]
        if self.size_offsets != 8 or self.size_lengths != 8:
            raise NotImplementedError("HDF5 feature not supported: %d-byte offsets" % self.size_offsets)
        pos = 24 + (4 if version == 1 else 0)  # v1 adds indexed-storage K + reserved
        pos += 32                              # base, free-space, EOF, driver-info addresses
        # root group symbol table entry: name offset, object header address, cache type, reserved, scratch
        _, root_addr, cache_type = struct.unpack_from("<QQI", self.buf, pos)
        self._cache = {}
        H5Group.__init__(self, self, root_addr)

    # ---- object headers -----------------------------------------------------------------------------
    def _read_object_header(self, addr):
        b = self.buf
        if b[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("HDF5 feature not supported: version-2 object headers")
        version, _, nmsgs, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise ValueError("bad object header version %d at %d" % (version, addr))
        msgs = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(msgs) < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:  # continuation
                    blocks.append(struct.unpack_from("<QQ", data, 0))
                msgs.append((mtype, data))
        return msgs

    def _open(self, addr):
        if addr not in self._cache:
            types = [t for t, _ in self._read_object_header(addr)]
            self._cache[addr] = H5Dataset(self, addr) if 0x0008 in types else H5Group(self, addr)
        return self._cache[addr]

    # ---- groups -------------------------------------------------------------------------------------
    def _heap_string(self, heap_addr, offset):
        b = self.buf
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("local heap signature missing at %d" % heap_addr)
        data_addr = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        start = data_addr + offset
        return b[start:b.index(b"\0", start)].decode("utf8")

    def _iter_group(self, btree, heap):
        b = self.buf
        sig = b[btree:btree + 4]
        if sig == b"SNOD":
            nsym = struct.unpack_from("<H", b, btree + 6)[0]
            for i in range(nsym):
                name_off, obj_addr = struct.unpack_from("<QQ", b, btree + 8 + 40 * i)
                yield self._heap_string(heap, name_off), obj_addr
            return
        if sig != b"TREE":
            raise ValueError("group B-tree signature missing at %d" % btree)
        node_type, _level, used = struct.unpack_from("<BBH", b, btree + 4)
        if node_type != 0:
            raise ValueError("expected a group B-tree node at %d" % btree)
        pos = btree + 24  # after signature, type, level, entries used, left + right sibling
        for i in range(used):
            child = struct.unpack_from("<Q", b, pos + 8 + 16 * i)[0]  # key_i (8), child_i (8), ...
            for item in self._iter_group(child, heap):
                yield item

    def _iter_chunks(self, btree, rank1):
        b = self.buf
        if b[btree:btree + 4] != b"TREE":
            raise ValueError("chunk B-tree signature missing at %d" % btree)
        node_type, level, used = struct.unpack_from("<BBH", b, btree + 4)
        if node_type != 1:
            raise ValueError("expected a chunk B-tree node at %d" % btree)
        key_size = 8 + 8 * rank1
        pos = btree + 24
        for i in range(used):
            kpos = pos + i * (key_size + 8)
            csize, fmask = struct.unpack_from("<II", b, kpos)
            offs = struct.unpack_from("<%dQ" % rank1, b, kpos + 8)[:-1]
            child = struct.unpack_from("<Q", b, kpos + key_size)[0]
            if fmask:
                raise NotImplementedError("HDF5 feature not supported: filtered chunks")
            if level > 0:
                for item in self._iter_chunks(child, rank1):
                    yield item
            else:
                yield offs, child, csize

    # ---- messages -----------------------------------------------------------------------------------
    @staticmethod
    def _parse_dataspace(d):
        version, rank, flags = d[0], d[1], d[2]
        if version == 1:
            pos = 8
        elif version == 2:
            pos = 4
        else:
            raise NotImplementedError("HDF5 feature not supported: dataspace message version %d" % version)
        return tuple(struct.unpack_from("<%dQ" % rank, d, pos)) if rank else ()

    @staticmethod
    def _parse_datatype(d):
        cls, version = d[0] & 0x0F, d[0] >> 4
        bits0 = d[1]
        size = struct.unpack_from("<I", d, 4)[0]
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:    # fixed point
            signed = bool(bits0 & 0x08)
            return _Datatype(cls, size, np.dtype("%s%s%d" % (order, "i" if signed else "u", size)))
        if cls == 1:    # floating point
            if size not in (2, 4, 8):
                raise NotImplementedError("HDF5 feature not supported: %d-byte floats" % size)
            return _Datatype(cls, size, np.dtype("%sf%d" % (order, size)))
        if cls == 3:    # fixed-length string
            return _Datatype(cls, size, np.dtype("S%d" % size), strpad=bits0 & 0x0F)
        if cls == 9:
            # variable length: bits 0-3 of the class bit field = 1 for strings (what h5py writes for Python str / bytes
            # attributes such as Keras' `backend` and `keras_version`); sequences of other types are not needed
            if (bits0 & 0x0F) != 1:
                raise NotImplementedError("HDF5 feature not supported: variable-length sequences")
            return _Datatype(cls, size, None, strpad=(bits0 >> 4) & 0x0F)
        raise NotImplementedError("HDF5 feature not supported: datatype class %d" % cls)

    def _attribute_name(self, d):
        version = d[0]
        name_size = struct.unpack_from("<H", d, 2)[0]
        pos = 8 + (1 if version == 3 else 0)
        return bytes(d[pos:pos + name_size]).split(b"\0")[0].decode("utf8", "replace")

    def _global_heap_object(self, addr, index):
        """Object `index` of the global heap collection at `addr` (HDF5 spec III.E: "GCOL", version, 3 reserved
        bytes, collection size; then objects {u16 index, u16 refcount, u32 reserved, length size, data padded to 8})."""
        b = self.buf
        if b[addr:addr + 4] != b"GCOL":
            raise ValueError("bad global heap collection at %d" % addr)
        total = struct.unpack_from("<Q", b, addr + 8)[0]
        pos, end = addr + 16, addr + total
        while pos + 16 <= end:
            idx, _, _, size = struct.unpack_from("<HHIQ", b, pos)
            if idx == 0:
                break  # free space marker
            if idx == index:
                return bytes(b[pos + 16:pos + 16 + size])
            pos += 16 + _pad8(size)
        raise ValueError("global heap object %d not found in the collection at %d" % (index, addr))

    def _parse_attribute(self, d):
        version = d[0]
        name_size, dt_size, ds_size = struct.unpack_from("<HHH", d, 2)
        if version == 1:
            pos = 8
            name = d[pos:pos + name_size]
            pos += _pad8(name_size)
            dt = d[pos:pos + dt_size]
            pos += _pad8(dt_size)
            ds = d[pos:pos + ds_size]
            pos += _pad8(ds_size)
        elif version in (2, 3):
            if d[1] & 0x03:
                raise NotImplementedError("HDF5 feature not supported: shared attribute datatypes")
            pos = 8 + (1 if version == 3 else 0)
            name = d[pos:pos + name_size]
            pos += name_size
            dt = d[pos:pos + dt_size]
            pos += dt_size
            ds = d[pos:pos + ds_size]
            pos += ds_size
        else:
            raise NotImplementedError("HDF5 feature not supported: attribute message version %d" % version)
        name = name.split(b"\0")[0].decode("utf8")
        dtype = self._parse_datatype(dt)
        shape = self._parse_dataspace(ds)
        n = int(np.prod(shape)) if len(shape) else 1
        if dtype.cls == 9:  # variable-length strings: {u32 length, heap collection address, u32 object index} each
            vals = []
            for k in range(n):
                length, gaddr, gidx = struct.unpack_from("<IQI", d, pos + 16 * k)
                vals.append(self._global_heap_object(gaddr, gidx)[:length] if length else b"")
            arr = np.asarray(vals, dtype=object).reshape(shape) if len(shape) else np.asarray(vals[0], dtype=object)
            return name, (arr if arr.shape else vals[0])
        arr = np.frombuffer(d, dtype=dtype.np_dtype, count=n, offset=pos).reshape(shape).copy()
        if dtype.cls == 3:
            arr = np.char.rstrip(arr, b"\0") if arr.shape else np.asarray(bytes(arr).rstrip(b"\0"))
        return name, (arr if arr.shape else arr[()])


# =====================================================================================================
# writer
# =====================================================================================================
class _Writer(object):
    """Appends HDF5 structures to a byte buffer; every structure starts 8-byte aligned."""

    GROUP_LEAF_K = 256     # a symbol-table node holds 2K entries: one SNOD per group suffices here
    GROUP_INTERNAL_K = 16

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- messages -----------------------------------------------------------------------------------
    @staticmethod
    def dataspace(shape):
        return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)

    @staticmethod
    def datatype(dt):
        dt = np.dtype(dt)
        if dt.kind == "f":
            spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
            sign, eloc, esize, mloc, msize, bias = spec
            # class 1, version 1; bit field: little endian, mantissa normalisation = implied MSB, sign position
            return struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dt.itemsize) + \
                struct.pack("<HHBBBBI", 0, dt.itemsize * 8, eloc, esize, mloc, msize, bias)
        if dt.kind in "iu":
            return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + \
                struct.pack("<HH", 0, dt.itemsize * 8)
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)  # null-padded ASCII
        raise TypeError("unsupported dtype %s" % dt)

    def vlen_string_attribute(self, name, value):
        """Scalar variable-length string attribute as h5py writes a Python str / bytes: the characters live in a global
        heap collection, the attribute holds {length, collection address, object index}."""
        raw = value if isinstance(value, bytes) else str(value).encode("utf8")
        obj = struct.pack("<HHIQ", 1, 1, 0, len(raw)) + raw + b"\0" * (_pad8(len(raw)) - len(raw))
        free = struct.pack("<HHIQ", 0, 0, 0, 0)
        total = 16 + len(obj) + len(free)
        gaddr = self.alloc(b"GCOL" + struct.pack("<B3xQ", 1, total) + obj + free)
        nm = name.encode("utf8") + b"\0"
        # datatype class 9 (variable length), version 1, type = string (1), padding null-terminated, charset 0;
        # size 16 (u32 + 8-byte address + u32); base type: 1-byte fixed string
        base = struct.pack("<BBBBI", 0x13, 0, 0, 0, 1)
        dt = struct.pack("<BBBBI", 0x19, 0x01, 0, 0, 16) + base
        ds = self.dataspace(())
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
        for part in (nm, dt, ds):
            body += part + b"\0" * (_pad8(len(part)) - len(part))
        return body + struct.pack("<IQI", len(raw), gaddr, 1)

    def attribute(self, name, value):
        if isinstance(value, VLenStr):
            return self.vlen_string_attribute(name, value.value)
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf8")
        if arr.dtype.kind == "S" and arr.dtype.itemsize == 0:
            arr = arr.astype("S1")
        nm = name.encode("utf8") + b"\0"
        dt, ds = self.datatype(arr.dtype), self.dataspace(arr.shape)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
        for part in (nm, dt, ds):
            body += part + b"\0" * (_pad8(len(part)) - len(part))
        return body + np.ascontiguousarray(arr).tobytes()

    def object_header(self, messages):
        body = b""
        for mtype, data in messages:
            data = data + b"\0" * (_pad8(len(data)) - len(data))
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        if len(body) > 0xFFFF0000:
            raise ValueError("object header too large")
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)

    # ---- objects ------------------------------------------------------------------------------------
    def dataset(self, arr, attrs=None):
        arr = np.asarray(arr)  # (np.ascontiguousarray would turn a scalar into shape (1,))
        data_addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        msgs = [(0x0001, self.dataspace(arr.shape)), (0x0003, self.datatype(arr.dtype)),
                (0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),                    # fill value v2: late alloc, undefined
                (0x0008, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes))]  # layout v3, contiguous
        for k, v in (attrs or {}).items():
            msgs.append((0x000C, self.attribute(k, v)))
        return self.object_header(msgs)

    def group(self, children, attrs=None):
        """children: {name: object header address}.  Returns (object header address, btree address, heap address)."""
        names = sorted(children)  # SNOD entries are ordered by name
        if len(names) > 2 * self.GROUP_LEAF_K:
            raise ValueError("too many links in one group (%d)" % len(names))
        heap_data = bytearray(b"\0" * 8)  # offset 0 = empty string
        offsets = {}
        for n in names:
            offsets[n] = len(heap_data)
            enc = n.encode("utf8") + b"\0"
            heap_data += enc + b"\0" * (_pad8(len(enc)) - len(enc))
        data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, data_addr))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:
            snod += struct.pack("<QQII16x", offsets[n], children[n], 0, 0)
        snod += b"\0" * (40 * (2 * self.GROUP_LEAF_K - len(names)))
        snod_addr = self.alloc(snod)
        # one leaf B-tree node: key0 = "" (offset 0), child0 = SNOD, key1 = largest name in the child
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        tree += struct.pack("<QQQ", 0, snod_addr, offsets[names[-1]] if names else 0)
        tree += b"\0" * (16 * (2 * self.GROUP_INTERNAL_K) + 8 - 24)
        btree_addr = self.alloc(tree)
        msgs = [(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        for k, v in (attrs or {}).items():
            msgs.append((0x000C, self.attribute(k, v)))
        return self.object_header(msgs), btree_addr, heap_addr

    def finish(self, root):
        root_addr, btree, heap = root
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack("<HHI", self.GROUP_LEAF_K, self.GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


def write_h5(path, tree, attrs=None):
    """tree: nested dicts; leaves are arrays, a dict value may carry its attributes under the key "__attrs__"."""
    w = _Writer()
    w.buf += b"\0" * 96  # superblock placeholder

    def build(node, node_attrs):
        children = {}
        for name, val in node.items():
            if name == "__attrs__":
                continue
            if isinstance(val, dict):
                children[name] = build(val, val.get("__attrs__"))[0]
            else:
                children[name] = w.dataset(np.asarray(val))
        return w.group(children, node_attrs)
    data = w.finish(build(tree, attrs))
    with open(path, "wb") as fh:
        fh.write(data)


# =====================================================================================================
# Keras weight files
# =====================================================================================================
def _short(weight_name):
    """b'encoder_L0_conv1/kernel:0' -> 'kernel'"""
    name = weight_name.decode("utf8") if isinstance(weight_name, bytes) else str(weight_name)
    return name.split("/")[-1].split(":")[0]


def load_keras_weights(path):
    """-> {layer name: {"kernel" | "bias" | "gamma" | "beta" | "moving_mean" | "moving_variance": array}} from a
    Keras weights file (or the `model_weights` group of a full-model file), like load_weights(by_name=True)."""
    f = H5File(path)
    root = f["model_weights"] if "layer_names" not in f.attrs and "model_weights" in f else f
    names = root.attrs.get("layer_names")
    if names is None:  # Keras splits long attributes into layer_names0, layer_names1, ...
        chunks = sorted((k for k in root.attrs if k.startswith("layer_names")), key=lambda k: int(k[11:] or 0))
        names = np.concatenate([np.atleast_1d(root.attrs[k]) for k in chunks]) if chunks else np.asarray(root.keys())
    out = {}
    for lname in np.atleast_1d(names):
        lname = lname.decode("utf8") if isinstance(lname, bytes) else str(lname)
        g = root[lname]
        wnames = g.attrs.get("weight_names")
        if wnames is None:
            chunks = sorted((k for k in g.attrs if k.startswith("weight_names")), key=lambda k: int(k[12:] or 0))
            wnames = np.concatenate([np.atleast_1d(g.attrs[k]) for k in chunks]) if chunks else []
        ws = {}
        for wn in np.atleast_1d(wnames):
            wn_s = wn.decode("utf8") if isinstance(wn, bytes) else str(wn)
            ws[_short(wn)] = g[wn_s].read()
        if ws:
            out[lname] = ws
    return out


def save_keras_weights(path, weights, layer_order=None, backend=b"tensorflow", keras_version=b"2.4.0"):
    """weights: {layer: {short name: array}} -> file laid out like Keras' save_weights (same group / dataset /
    attribute names), float32 datasets."""
    order = list(layer_order or weights.keys())
    per_layer_order = ("kernel", "bias", "gamma", "beta", "moving_mean", "moving_variance")
    tree = {}
    for lname in order:
        ws = weights[lname]
        keys = [k for k in per_layer_order if k in ws] + [k for k in ws if k not in per_layer_order]
        inner = {"%s:0" % k: np.asarray(ws[k], dtype=np.float32) for k in keys}
        wnames = np.asarray([("%s/%s:0" % (lname, k)).encode("utf8") for k in keys])
        tree[lname] = {"__attrs__": {"weight_names": wnames}, lname: inner}
    attrs = {"layer_names": np.asarray([n.encode("utf8") for n in order]),
             "backend": np.asarray(backend), "keras_version": np.asarray(keras_version)}
    write_h5(path, tree, attrs)
