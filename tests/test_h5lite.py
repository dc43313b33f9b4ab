"""CPU: the package's own HDF5 reader / writer (bsc-nav_b200/h5lite.py), the file family h5py writes for the reference's
feat.h5df (memory_2.py:335-354).  libhdf5 is not in this image, so the reader is pinned on a file libhdf5 itself wrote -- scipy
ships one with its test data (MATLAB v7.3 = HDF5 behind a 512-byte user block; expected values from scipy's own test_mio.py) -- and
the writer on the reader plus structural checks against the published format."""
import importlib
import os
import struct

import numpy as np
import pytest

h5 = importlib.import_module("bsc-nav_b200.h5lite")


def test_reads_a_file_written_by_libhdf5():
    import scipy.io.matlab
    path = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy test data not installed")
    with h5.File(path, "r") as f:
        assert f.keys() == ["testdouble"] and "testdouble" in f and "nope" not in f
        ds = f["testdouble"]
        assert ds.shape == (9, 1) and ds.dtype == np.dtype("<f8")
        np.testing.assert_array_equal(ds[:].ravel(), np.pi / 4 * np.arange(9))       # scipy/io/matlab/tests/test_mio.py:49-50


def test_token_file_round_trip_with_deep_group_tree(tmp_path):
    """The reference's layout at a size that needs a three-level group B-tree (2 500 voxels > 8 x 32 x 8)."""
    rng = np.random.default_rng(0)
    p = str(tmp_path / "feat.h5df")
    want = {}
    with h5.File(p, "w") as f:
        while len(want) < 2500:
            name = f"grid_{rng.integers(0, 512)}_{rng.integers(0, 512)}_{rng.integers(0, 200)}"
            if name in f:
                continue
            m = int(rng.integers(1, 11))
            feats, dist = rng.standard_normal((m, 32)).astype(np.float32), rng.random(m).astype(np.float32)
            g = f.create_group(name)
            g.create_dataset("features", data=feats, maxshape=(None, feats.shape[1]), chunks=True)
            g.create_dataset("distances", data=dist, maxshape=(None,), chunks=True)
            want[name] = (feats, dist)
        with pytest.raises(ValueError):
            f.create_group(next(iter(want)))                 # h5py raises the same way on an existing name
    with h5.File(p, "r") as f:
        assert f.keys() == sorted(want, key=lambda s: s.encode())       # HDF5 iterates old-style groups in name order
        for n, (feats, dist) in want.items():
            assert f[n].keys() == ["distances", "features"]
            assert f[n]["features"].shape == feats.shape and f[n]["features"].maxshape == (None, 32)
            assert np.array_equal(f[n]["features"][:], feats) and np.array_equal(f[n]["distances"][:], dist)
            assert f[n]["distances"][-1] == dist[-1] and f[n]["features"][0:1].shape == (1, 32)


def test_written_structures_follow_the_format(tmp_path):
    p = str(tmp_path / "s.h5")
    with h5.File(p, "w") as f:
        g = f.create_group("grid_1_2_3")
        g.create_dataset("features", data=np.arange(6, dtype=np.float32).reshape(2, 3), maxshape=(None, 3), chunks=True)
        f.create_dataset("fixed", data=np.arange(5, dtype=np.int64))
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8          # superblock 0, 8-byte offsets / lengths
    leaf_k, node_k = struct.unpack_from("<HH", b, 16)
    assert (leaf_k, node_k) == (4, 16)
    eof = struct.unpack_from("<Q", b, 40)[0]
    assert eof == len(b)
    root_hdr, cache_type = struct.unpack_from("<Q", b, 64)[0], struct.unpack_from("<I", b, 72)[0]
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert cache_type == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP" and b[root_hdr] == 1
    # root header: one symbol-table message (0x0011) pointing at the same B-tree and heap
    n_msg = struct.unpack_from("<H", b, root_hdr + 2)[0]
    mtype, msize = struct.unpack_from("<HH", b, root_hdr + 16)
    assert n_msg == 1 and mtype == 0x0011 and struct.unpack_from("<QQ", b, root_hdr + 24) == (btree, heap) and msize == 16
    leaf = struct.unpack_from("<Q", b, btree + 32)[0]
    assert b[leaf:leaf + 4] == b"SNOD" and struct.unpack_from("<H", b, leaf + 6)[0] == 2
    # heap: free-list head inside the data segment, names null-terminated and 8-byte aligned
    size, free, data = struct.unpack_from("<QQQ", b, heap + 8)
    seg = b[data:data + size]
    assert free < size and seg[:8] == bytes(8) and b"fixed\0" in seg and b"grid_1_2_3\0" in seg
    with h5.File(p, "r") as f:
        assert np.array_equal(f["fixed"][:], np.arange(5)) and f["fixed"].dtype == np.dtype("<i8")
        assert f["grid_1_2_3"]["features"].maxshape == (None, 3)


def test_rejects_what_it_does_not_understand(tmp_path):
    p = str(tmp_path / "x.h5")
    open(p, "wb").write(b"not hdf5" * 200)
    with pytest.raises(h5.H5Error):
        h5.File(p, "r")
    sb2 = bytearray(b"\x89HDF\r\n\x1a\n" + bytes(200))
    sb2[8] = 2                                              # superblock version 2: libver='latest' files
    open(p, "wb").write(bytes(sb2))
    with pytest.raises(h5.H5Error, match="superblock version 2"):
        h5.File(p, "r")
    with pytest.raises(h5.H5Error):
        h5.File(p, "a")
