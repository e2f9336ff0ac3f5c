"""File formats at the edges of the path: the native parsers (keystone_b200/csrc/io.cu) against numpy on
files written here -- headerless CSV (K/loaders/CsvDataLoader.scala:28-30), TIMIT sparse labels
(K/loaders/TimitFeaturesDataLoader.scala:22-42), CIFAR-10 records (K/loaders/CifarLoader.scala:30-45).  Host code: runs without a GPU."""
import numpy as np
import pytest

import keystone_b200 as ks


def test_csv_loader_matches_numpy(tmp_path):
    rng = np.random.default_rng(0)
    m = rng.standard_normal((1000, 37)) * 10.0 ** rng.integers(-8, 8, (1000, 37))
    p = tmp_path / "a.csv"
    np.savetxt(p, m, delimiter=",", fmt="%.17g")
    got = ks.CsvDataLoader(str(p), np.float64)
    assert got.shape == m.shape and np.array_equal(got, np.loadtxt(p, delimiter=","))
    got32 = ks.CsvDataLoader(str(p))
    assert got32.dtype == np.float32 and np.array_equal(got32, m.astype(np.float32))
    # the reference's own fixture format: no header, no trailing newline needed, CRLF tolerated
    (tmp_path / "b.csv").write_bytes(b"1,2.5,-3e2\r\n4,5,6")
    assert np.array_equal(ks.CsvDataLoader(str(tmp_path / "b.csv"), np.float64), np.array([[1, 2.5, -300.0], [4, 5, 6]]))


def test_csv_loader_reports_ragged_and_bad_rows(tmp_path):
    (tmp_path / "r.csv").write_text("1,2,3\n4,5\n")
    with pytest.raises(ks.KeystoneError):
        ks.CsvDataLoader(str(tmp_path / "r.csv"))
    (tmp_path / "x.csv").write_text("1,abc,3\n")
    with pytest.raises(ks.KeystoneError):
        ks.CsvDataLoader(str(tmp_path / "x.csv"))
    with pytest.raises(ks.KeystoneError):
        ks.CsvDataLoader(str(tmp_path / "missing.csv"))


def test_mnist_csv_layout(tmp_path):
    """MnistRandomFFT.scala:34-36: label = x(0).toInt - 1, data = x(1 until x.length)."""
    rng = np.random.default_rng(1)
    lab = rng.integers(1, 11, 50)
    pix = rng.integers(0, 256, (50, 784))
    np.savetxt(tmp_path / "m.csv", np.column_stack([lab, pix]), delimiter=",", fmt="%d")
    d = ks.MnistCsvLoader(str(tmp_path / "m.csv"))
    assert np.array_equal(d.labels, lab - 1) and np.array_equal(d.data, pix.astype(np.float32))


def test_timit_labels(tmp_path):
    """TimitFeaturesDataLoader.scala:26-42: 'row label' lines, both 1-based, any order."""
    rng = np.random.default_rng(2)
    n = 300
    lab = rng.integers(1, 148, n)
    order = rng.permutation(n)
    (tmp_path / "l.txt").write_text("".join(f"{r + 1} {lab[r]}\n" for r in order))
    assert np.array_equal(ks.TimitLabelsLoader(str(tmp_path / "l.txt"), n), lab - 1)
    (tmp_path / "short.txt").write_text("1 5\n")
    with pytest.raises(ks.KeystoneError):
        ks.TimitLabelsLoader(str(tmp_path / "short.txt"), 2)


def test_cifar_records(tmp_path):
    """CifarLoader.scala:30-45: 1 label byte + 3072 image bytes per record."""
    rng = np.random.default_rng(3)
    n = 20
    lab = rng.integers(0, 10, n).astype(np.uint8)
    img = rng.integers(0, 256, (n, 3072)).astype(np.uint8)
    (tmp_path / "c.bin").write_bytes(b"".join(bytes([lab[i]]) + img[i].tobytes() for i in range(n)))
    d = ks.CifarLoader(str(tmp_path / "c.bin"))
    assert np.array_equal(d.labels, lab.astype(np.int32)) and np.array_equal(d.data.reshape(n, 3072), img)
    (tmp_path / "bad.bin").write_bytes(b"\x00" * 100)
    with pytest.raises(ks.KeystoneError):
        ks.CifarLoader(str(tmp_path / "bad.bin"))
