"""utils/keras_h5.py: the HDF5 subset reader/writer for Keras weight files.  No h5py exists in the build
environment, so these tests pin (a) round trips through the package's own writer, (b) the byte layout of the
structures against the offsets the HDF5 file-format specification prescribes (superblock fields, object-header
prefix, heap / B-tree / SNOD signatures), (c) the reader on hand-assembled variants the writer never emits
(compact and chunked layouts, version-2 dataspace, version-3 attribute, object-header continuation), and
(d) loud errors for unsupported features."""
import struct

import numpy as np
import pytest

from multiplanarunet_b200.utils import keras_h5 as K


def _weights(rng):
    return {
        "encoder_L0_conv1": {"kernel": rng.randn(3, 3, 1, 6).astype(np.float32), "bias": rng.randn(6).astype(np.float32)},
        "encoder_L0_BN": {"gamma": rng.rand(6).astype(np.float32), "beta": rng.randn(6).astype(np.float32),
                          "moving_mean": rng.randn(6).astype(np.float32),
                          "moving_variance": rng.rand(6).astype(np.float32)},
        "conv2d": {"kernel": rng.randn(1, 1, 6, 3).astype(np.float32), "bias": np.zeros(3, np.float32)},
    }


def test_keras_weight_file_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    w = _weights(rng)
    path = str(tmp_path / "model_weights.h5")
    K.save_keras_weights(path, w, layer_order=["encoder_L0_conv1", "encoder_L0_BN", "conv2d"])
    back = K.load_keras_weights(path)
    assert list(back) == ["encoder_L0_conv1", "encoder_L0_BN", "conv2d"]      # layer_names order
    for layer in w:
        assert set(back[layer]) == set(w[layer])
        for k in w[layer]:
            assert back[layer][k].dtype == np.float32 and np.array_equal(back[layer][k], w[layer][k])
    f = K.H5File(path)
    assert f.attrs["backend"] == b"tensorflow" and f.attrs["keras_version"] == b"2.4.0"
    assert list(f.attrs["layer_names"]) == [b"encoder_L0_conv1", b"encoder_L0_BN", b"conv2d"]
    g = f["encoder_L0_BN"]
    assert list(g.attrs["weight_names"]) == [b"encoder_L0_BN/gamma:0", b"encoder_L0_BN/beta:0",
                                             b"encoder_L0_BN/moving_mean:0", b"encoder_L0_BN/moving_variance:0"]
    d = f["encoder_L0_conv1/encoder_L0_conv1/kernel:0"]
    assert d.shape == (3, 3, 1, 6) and d.dtype == np.dtype("<f4")
    with pytest.raises(KeyError):
        f["nope/kernel:0"]
    # a full-model file keeps the same tree under /model_weights
    tree = {"model_weights": {"__attrs__": {"layer_names": np.asarray([b"conv2d"])},
                              "conv2d": {"__attrs__": {"weight_names": np.asarray([b"conv2d/kernel:0", b"conv2d/bias:0"])},
                                         "conv2d": {"kernel:0": w["conv2d"]["kernel"], "bias:0": w["conv2d"]["bias"]}}}}
    K.write_h5(str(tmp_path / "full.h5"), tree)
    full = K.load_keras_weights(str(tmp_path / "full.h5"))
    assert np.array_equal(full["conv2d"]["kernel"], w["conv2d"]["kernel"])


def test_layout_follows_the_format_specification(tmp_path):
    path = str(tmp_path / "t.h5")
    K.write_h5(path, {"g": {"x": np.arange(6, dtype=np.float32).reshape(2, 3)}, "i": np.arange(4, dtype=np.int32)},
               attrs={"note": np.asarray(b"hi")})
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0              # superblock version 0
    assert b[13] == 8 and b[14] == 8                                  # size of offsets / lengths
    leaf_k, internal_k = struct.unpack_from("<HH", b, 16)
    base, free, eof, driver = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == K.UNDEF and driver == K.UNDEF and eof == len(b)
    name_off, root, cache_type = struct.unpack_from("<QQI", b, 56)
    btree, heap = struct.unpack_from("<QQ", b, 80)                   # scratch-pad of a cached group entry
    assert cache_type == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    # object header prefix: version 1, message count, reference count 1, header size; messages 8-aligned
    ver, _, nmsg, refs, hsize = struct.unpack_from("<BBHII", b, root)
    assert ver == 1 and refs == 1 and nmsg == 2 and root % 8 == 0 and hsize % 8 == 0
    mtype, msize = struct.unpack_from("<HH", b, root + 16)
    assert mtype == 0x11 and msize == 16 and struct.unpack_from("<QQ", b, root + 24) == (btree, heap)
    # the B-tree leaf points at one SNOD whose entries are sorted by name
    node_type, level, used = struct.unpack_from("<BBH", b, btree + 4)
    snod = struct.unpack_from("<Q", b, btree + 24 + 8)[0]
    assert (node_type, level, used) == (0, 0, 1) and b[snod:snod + 4] == b"SNOD"
    assert struct.unpack_from("<H", b, snod + 6)[0] == 2
    assert len(b[snod:]) >= 8 + 2 * leaf_k * 40                       # full-size symbol node
    f = K.H5File(path)
    assert f.keys() == ["g", "i"] and f.attrs["note"] == b"hi"
    assert np.array_equal(f["g/x"].read(), np.arange(6, dtype=np.float32).reshape(2, 3))
    assert f["i"].read().dtype == np.dtype("<i4") and list(f["i"].read()) == [0, 1, 2, 3]


def test_reader_handles_layouts_and_message_versions_the_writer_does_not_emit(tmp_path):
    w = K._Writer()
    w.buf += b"\0" * 96
    data = np.arange(12, dtype=np.float32).reshape(3, 4)
    # compact layout + version-2 dataspace
    ds_v2 = struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 3, 4)
    compact = w.object_header([(0x0001, ds_v2), (0x0003, w.datatype(np.float32)),
                               (0x0008, struct.pack("<BBH", 3, 0, data.nbytes) + data.tobytes())])
    # chunked layout (2x4 chunks, second chunk partly outside the array), no filters
    c0 = w.alloc(data[0:2].tobytes())
    c1 = w.alloc(np.vstack([data[2:3], np.full((1, 4), -1, np.float32)]).tobytes())
    keys = [(32, 0, (0, 0, 0)), (32, 0, (2, 0, 0)), (0, 0, (4, 0, 0))]
    tree = b"TREE" + struct.pack("<BBHQQ", 1, 0, 2, K.UNDEF, K.UNDEF)
    for (size, mask, offs), child in zip(keys, [c0, c1, None]):
        tree += struct.pack("<II3Q", size, mask, *offs)
        if child is not None:
            tree += struct.pack("<Q", child)
    bt = w.alloc(tree)
    # the version-3 attribute and the layout message sit in a continuation block
    attr_v3 = struct.pack("<BBHHHB", 3, 0, 5, 8, len(w.dataspace(())), 0) + b"unit\0" + w.datatype(np.dtype("S2")) + \
        w.dataspace(()) + b"mm"
    hdr = w.object_header([(0x0001, w.dataspace((3, 4))), (0x0003, w.datatype(np.float32))])
    cbody = b""
    for mtype, d in [(0x0008, struct.pack("<BBBQ3I", 3, 2, 3, bt, 2, 4, 4)), (0x000C, attr_v3)]:
        d = d + b"\0" * (K._pad8(len(d)) - len(d))
        cbody += struct.pack("<HHB3x", mtype, len(d), 0) + d
    caddr = w.alloc(cbody)
    chunked = w.object_header([(0x0001, w.dataspace((3, 4))), (0x0003, w.datatype(np.float32)),
                               (0x0010, struct.pack("<QQ", caddr, len(cbody)))])
    # patch the message count: the continuation block carries two more messages
    struct.pack_into("<H", w.buf, chunked + 2, 5)
    raw = w.finish(w.group({"compact": compact, "chunked": chunked}))
    path = str(tmp_path / "v.h5")
    open(path, "wb").write(raw)
    f = K.H5File(path)
    assert np.array_equal(f["compact"].read(), data)
    assert np.array_equal(f["chunked"].read(), data)
    assert f["chunked"].attrs["unit"] == b"mm"


def test_unsupported_features_fail_loudly(tmp_path):
    p = str(tmp_path / "x.h5")
    open(p, "wb").write(b"not an hdf5 file at all")
    with pytest.raises(ValueError):
        K.H5File(p)
    w = K._Writer()
    w.buf += b"\0" * 96
    raw = bytearray(w.finish(w.group({})))
    raw[8] = 2                                      # superblock version 2
    open(p, "wb").write(bytes(raw))
    with pytest.raises(NotImplementedError):
        K.H5File(p)
    # filtered dataset (filter pipeline message present)
    w = K._Writer()
    w.buf += b"\0" * 96
    d = w.object_header([(0x0001, w.dataspace((2,))), (0x0003, w.datatype(np.float32)),
                         (0x000B, b"\x01\x01" + b"\0" * 6), (0x0008, struct.pack("<BBQQ", 3, 1, K.UNDEF, 8))])
    open(p, "wb").write(w.finish(w.group({"z": d})))
    with pytest.raises(NotImplementedError):
        K.H5File(p)["z"]
    with pytest.raises(NotImplementedError):
        K.H5File._parse_datatype(struct.pack("<BBBBI", 0x19, 0, 0, 0, 16))      # variable-length string


def test_variable_length_string_attributes_and_unsupported_attributes_do_not_block_loading(tmp_path):
    """Keras / h5py store `backend` and `keras_version` on the root group as VARIABLE-LENGTH strings (global heap
    objects).  The reader decodes them, and an attribute of a type it cannot decode only fails when that attribute is
    read - never when the file is opened or the weights are loaded.  (Still self-validation: no h5py here.)"""
    import struct
    from multiplanarunet_b200.utils import keras_h5 as K
    tree = {"layer_a": {"__attrs__": {"weight_names": np.array([b"layer_a/kernel:0"])},
                        "layer_a": {"kernel:0": np.arange(6, dtype=np.float32).reshape(2, 3)}}}
    attrs = {"layer_names": np.array([b"layer_a"]), "backend": K.VLenStr("tensorflow"),
             "keras_version": K.VLenStr(b"2.4.0")}
    path = str(tmp_path / "vlen.h5")
    K.write_h5(path, tree, attrs)
    f = K.H5File(path)
    assert f.attrs["backend"] == b"tensorflow" and f.attrs["keras_version"] == b"2.4.0"
    w = K.load_keras_weights(path)
    assert np.array_equal(w["layer_a"]["kernel"], np.arange(6, dtype=np.float32).reshape(2, 3))
    # corrupt the vlen datatype into a variable-length SEQUENCE (class 9, type 0): unsupported -> lazy error
    raw = bytearray(open(path, "rb").read())
    pat = struct.pack("<BBBBI", 0x19, 0x01, 0, 0, 16)
    at = raw.find(pat)
    assert at > 0
    raw[at + 1] = 0x00
    bad = str(tmp_path / "seq.h5")
    open(bad, "wb").write(bytes(raw))
    g = K.H5File(bad)                               # opens
    assert "layer_a" in K.load_keras_weights(bad)   # loads
    with pytest.raises(NotImplementedError):
        g.attrs["backend"]
