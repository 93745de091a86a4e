"""minih5: the saved-weights files are real HDF5 (superblock v0, symbol-table groups, contiguous datasets).
Round trip through the library's own reader, byte-level checks of the structures a foreign reader walks, and a
cross-read with h5py wherever it is installed."""
import struct

import numpy as np
import pytest

from anime_recommendations_b200 import minih5


def _tree():
    g = minih5.Group(attrs=dict(keras_version=b"2.12.0", backend=b"tensorflow", model_config=b'{"class_name": "Functional"}'))
    mw = g.group("model_weights")
    mw.attrs["layer_names"] = [b"user", b"user_embedding", b"dense"]
    mw.group("user").attrs["weight_names"] = []
    mw.group("user_embedding").attrs["weight_names"] = [b"user_embedding/embeddings:0"]
    rng = np.random.RandomState(0)
    mw.dataset("user_embedding/user_embedding/embeddings:0", rng.standard_normal((37, 16)).astype(np.float32))
    mw.dataset("dense/dense/kernel:0", np.array([[1.25]], np.float32))
    mw.dataset("dense/dense/bias:0", np.zeros(1, np.float32))
    g.dataset("optimizer_weights/iteration:0", np.array(1234, np.int64))
    for i in range(19):                                    # more members than one symbol-table node holds
        g.dataset("wide/m%02d" % i, np.full((2, 3), i, np.float64), attrs=dict(idx=np.int32(i)))
    return g


def test_round_trip(tmp_path):
    p = str(tmp_path / "t.h5")
    minih5.write(p, _tree())
    d, a = minih5.read(p)
    assert a[""]["keras_version"].tobytes() == b"2.12.0" and a[""]["backend"].tobytes() == b"tensorflow"
    assert a["/model_weights"]["layer_names"].tolist() == [b"user", b"user_embedding", b"dense"]
    assert a["/model_weights/user"]["weight_names"].shape == (0,)
    assert a["/model_weights/user_embedding"]["weight_names"].tolist() == [b"user_embedding/embeddings:0"]
    e = d["/model_weights/user_embedding/user_embedding/embeddings:0"]
    np.testing.assert_array_equal(e, np.random.RandomState(0).standard_normal((37, 16)).astype(np.float32))
    assert d["/model_weights/dense/dense/kernel:0"].shape == (1, 1) and d["/optimizer_weights/iteration:0"] == 1234
    assert sorted(k for k in d if k.startswith("/wide/")) == ["/wide/m%02d" % i for i in range(19)]
    assert all(int(a["/wide/m%02d" % i]["idx"]) == i and d["/wide/m%02d" % i][1, 2] == i for i in range(19))


def test_file_structure_bytes(tmp_path):
    """What libhdf5 checks first: signature, superblock v0 fields, end-of-file address, root symbol-table entry ->
    object header v1 -> symbol-table message -> B-tree 'TREE' / heap 'HEAP' / 'SNOD' with sorted names."""
    p = str(tmp_path / "t.h5")
    minih5.write(p, _tree())
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])                   # versions 0, 8-byte offsets and lengths
    leaf_k, internal_k = struct.unpack_from("<HH", b, 16)
    assert (leaf_k, internal_k) == (4, 16)
    base, freesp, eof, driver = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and freesp == driver == 0xFFFFFFFFFFFFFFFF and eof == len(b)
    name_off, hdr, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert name_off == 0 and cache == 1
    assert b[hdr] == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    nmsg, refs, size = struct.unpack_from("<HII", b, hdr + 2)
    assert refs == 1 and size % 8 == 0
    mtype, msize = struct.unpack_from("<HH", b, hdr + 16)
    assert mtype == 0x11 and struct.unpack_from("<QQ", b, hdr + 24) == (btree, heap)
    used = struct.unpack_from("<H", b, btree + 6)[0]
    seg = struct.unpack_from("<Q", b, heap + 24)[0]
    names = []
    for i in range(used):
        snod = struct.unpack_from("<Q", b, btree + 32 + 16 * i)[0]
        assert b[snod:snod + 4] == b"SNOD"
        for j in range(struct.unpack_from("<H", b, snod + 6)[0]):
            off = struct.unpack_from("<Q", b, snod + 8 + 40 * j)[0]
            names.append(b[seg + off:b.index(b"\x00", seg + off)])
    assert names == sorted(names) == [b"model_weights", b"optimizer_weights", b"wide"]
    assert b.count(b"user_embedding/embeddings:0") >= 1


def test_unsupported_dtype_and_oversized_attribute_raise(tmp_path):
    g = minih5.Group()
    g.dataset("x", np.zeros(3, np.float16))
    with pytest.raises(TypeError):
        minih5.write(str(tmp_path / "a.h5"), g)
    g = minih5.Group(attrs=dict(big=b"x" * 70000))
    with pytest.raises(ValueError):
        minih5.write(str(tmp_path / "b.h5"), g)


def test_h5py_reads_the_file(tmp_path):
    h5py = pytest.importorskip("h5py")
    p = str(tmp_path / "t.h5")
    minih5.write(p, _tree())
    with h5py.File(p, "r") as f:
        assert f.attrs["keras_version"] == b"2.12.0"
        assert list(f["model_weights"].attrs["layer_names"]) == [b"user", b"user_embedding", b"dense"]
        np.testing.assert_array_equal(f["model_weights/dense/dense/kernel:0"][()], [[1.25]])
        assert len(f["wide"]) == 19 and int(f["optimizer_weights/iteration:0"][()]) == 1234


def test_reads_a_keras_written_file():
    """tests/golden/tf_model_keras.h5 is what Keras 2.12 + h5py write for the reference's model (produced by
    tests/golden/make_tf_golden.py wherever TensorFlow exists); the reader must find the Keras tensor names in it."""
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_model_keras.h5")
    if not os.path.exists(p):
        pytest.skip("no Keras-written file (run tests/golden/make_tf_golden.py where TF 2.12 exists)")
    d, a = minih5.read(p)
    assert d["/model_weights/user_embedding/user_embedding/embeddings:0"].ndim == 2
    assert [x.decode() for x in a["/model_weights"]["layer_names"].tolist()][:2] == ["user", "anime"]
