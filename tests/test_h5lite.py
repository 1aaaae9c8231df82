"""CPU: the built-in HDF5 writer / reader (mara3_b200/csrc/h5lite.cpp; libhdf5 is not in this image).

1. tests/h5_reader.py -- an independent parser of the on-disk format -- is pinned against a file written by the REAL
   library (tests/golden/libhdf5_sample_v73.mat: scipy's MATLAB v7.3 test file, libhdf5 1.6 behind MATLAB 7.4;
   version-0 superblock behind a 512-byte user block, symbol-table root group, one contiguous f64 dataset).
2. The file the C++ writer produces is parsed with that reader: groups of 0 / 1 / 300 entries (one and two B-tree
   levels), scalar / array / compound / string datatypes, an empty (0, 0) dataset.
3. The C++ reader lists both files."""
import ctypes as C
import os
import struct
import numpy as np
import pytest
import mara3_b200 as m3
from h5_reader import H5File, SIGNATURE

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "libhdf5_sample_v73.mat")


def selftest(write_path="", read_path=""):
    lib = m3.load_library()
    lib.m3b_h5_selftest.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(4096)
    rc = lib.m3b_h5_selftest(write_path.encode(), read_path.encode(), buf, 4096)
    return rc, buf.value.decode()


def test_python_reader_reads_a_file_written_by_libhdf5():
    f = H5File(GOLDEN)
    assert f.base == 512 and f.superblock_version == 0 and (f.leaf_k, f.internal_k) == (4, 16)
    assert f.keys("/") == ["testdouble"]
    assert f.shape("/testdouble") == (9, 1) and f.dtype("/testdouble") == np.dtype("<f8")
    assert np.allclose(f.read("/testdouble").ravel(), np.pi * np.arange(9) / 4, rtol=0, atol=1e-15)


def test_cpp_reader_reads_a_file_written_by_libhdf5():
    rc, report = selftest(read_path=GOLDEN)
    assert rc == 0 and report.split() == ["testdouble"], report


def test_writer_output_parses_as_hdf5(tmp_path):
    path = str(tmp_path / "selftest.h5")
    rc, report = selftest(write_path=path, read_path=path)
    assert rc == 0, report
    assert report.split() == ["group/", "hollow/", "many/", "series"]

    raw = open(path, "rb").read()
    assert raw[:8] == SIGNATURE and struct.unpack_from("<Q", raw, 40)[0] == len(raw)       # end-of-file address
    f = H5File(path)
    assert f.keys("/") == ["group", "hollow", "many", "series"]
    assert f.keys("/hollow") == []
    assert f.keys("/group") == ["count", "empty", "iteration", "name", "nested", "time"]
    assert f.read("/group/time") == 3.25 and f.dtype("/group/time") == np.dtype("<f8") and f.shape("/group/time") == ()
    assert f.read("/group/count") == -7 and f.dtype("/group/count") == np.dtype("<i4")
    assert f.read("/group/name") == b"write_checkpoint" and f.dtype("/group/name") == np.dtype("S16")
    assert f.dtype("/group/empty") == np.dtype("S1") and f.read("/group/empty") == b""
    assert f.dtype("/group/iteration") == np.dtype(("<i4", (2,))) and list(f.read("/group/iteration")) == [22, 7]

    field = f.read("/group/nested/field")
    assert f.shape("/group/nested/field") == (6, 4) and field.shape == (6, 4, 3)
    assert np.array_equal(field.ravel(), 0.125 * np.arange(72))
    assert f.shape("/group/nested/nothing") == (0, 0) and f.layout("/group/nested/nothing")[1] == 0xFFFFFFFFFFFFFFFF

    series = f.read("/series")
    dt = f.dtype("/series")
    assert dt.names == ("t", "v", "in", "n") and dt.itemsize == 48
    assert [dt.fields[n][1] for n in dt.names] == [0, 8, 24, 40]
    assert dt.fields["in"][0].names == ("a", "b")
    assert np.array_equal(series["t"], 0.5 * np.arange(5)) and np.array_equal(series["v"][:, 1], 2.0 + np.arange(5))
    assert np.array_equal(series["in"]["b"], -1.0 * np.arange(5)) and np.array_equal(series["n"], np.arange(5))

    # 300 entries: 38 symbol-table nodes under two level-0 B-tree nodes under a level-1 root
    names = f.keys("/many")
    assert len(names) == 300 and names == sorted("8:%03d-%03d" % (k, 299 - k) for k in range(300))
    for k in (0, 7, 8, 255, 256, 299):
        assert f.read("/many/8:%03d-%03d" % (k, 299 - k))[0] == float(k)
    for mtype, flags, data in f.messages(f.lookup("/many")):
        if mtype == 0x0011:
            root = struct.unpack_from("<Q", data, 0)[0]
            assert f.at(root, 4) == b"TREE" and f.at(root, 8)[5] == 1 and struct.unpack_from("<H", f.at(root, 8), 6)[0] == 2


def test_writer_refuses_duplicates_and_reader_reports_missing(tmp_path):
    rc, report = selftest(read_path=str(tmp_path / "absent.h5"))
    assert rc == -1 and "cannot open" in report


def test_leaf_dataset_names_match_the_reference_known_answers():
    """mara::format_tree_index known answers (Mara3 src/app_test.cpp:379-382), 2-D form: zero padded to the width of 2^level."""
    lib = m3.load_library()
    lib.m3b_format_tree_index.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    def name(level, i, j):
        buf = C.create_string_buffer(64)
        lib.m3b_format_tree_index(level, i, j, buf, 64)
        return buf.value.decode()
    assert name(0, 0, 0) == "0:0-0"
    assert name(3, 5, 6) == "3:5-6"
    assert name(5, 1, 16) == "5:01-16"
    assert name(8, 1, 2) == "8:001-002"
    assert name(4, 15, 0) == "4:15-00" and name(10, 1023, 7) == "10:1023-0007"


def test_a_file_shared_between_ranks_is_the_file_one_rank_writes(tmp_path):
    """Multi-GPU product writers: rank 0 writes the structure and its own blocks and leaves holes, every other rank stores its
    blocks at the addresses it has computed itself (h5lite roles root / part).  Byte for byte the single-writer file."""
    lib = m3.load_library()
    lib.m3b_h5_selftest_shared.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    for parts in (2, 3, 8):
        whole, shared = tmp_path / f"whole{parts}.h5", tmp_path / f"shared{parts}.h5"
        assert lib.m3b_h5_selftest_shared(str(whole).encode(), str(shared).encode(), parts) == 0
        a, b = whole.read_bytes(), shared.read_bytes()
        assert len(a) == len(b) and a == b
    f = H5File(str(tmp_path / "shared8.h5"))
    assert len(f.keys("/solution/conserved_u")) == 37
    block = f.read("/solution/conserved_u/6:36-01")
    assert block.shape == (5, 5, 3) and block[0, 0, 0] == 1.0 / (36 * 75 + 1)
