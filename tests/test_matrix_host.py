"""CPU test of digest.build_matrix (the reference's matrix build, mirge/libs/digest.py:237-261) with a stand-in for
the device table: DataFrame contract of SURVEY.md section 8b -- index name, column order, dtypes, '' fill, row order,
keys no sample counted are left out -- and the pickle round trip of `-spl` / `-rr` (__main__.py:101,145)."""
import pickle

import numpy as np
import pandas as pd

import mirge_b200  # noqa: F401
from mirge_b200 import digest as DG


class FakeTable:
    def __init__(self, keys):
        w = max(len(k) for k in keys)
        self._keys = np.array([k.encode() for k in keys], dtype="S%d" % w)

    @property
    def n_keys(self):
        return int(self._keys.shape[0])

    def export_keys(self):
        return self._keys

    def export_sorted(self, seen=None):
        """host stand-in of CollapseTable.export_sorted: order, offsets, packed texts"""
        order = np.argsort(self._keys, kind="stable")
        if seen is not None:
            order = order[seen[order]]
        texts = [bytes(k) for k in self._keys[order].tolist()]
        offsets = np.zeros(len(texts) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(t) for t in texts])
        return order, offsets, np.frombuffer(b"".join(texts), dtype=np.uint8)


def sample(ids, counts):
    r = DG.SampleResult()
    r.ids = np.array(ids, dtype=np.int64)
    r.counts = np.array(counts, dtype=np.int64)
    r.count, r.trimmed, r.unique = int(sum(counts)), int(sum(counts)), len(ids)
    r.hist = np.zeros(0, dtype=np.int64)
    return r


def test_matrix_contract_and_pickle_round_trip(tmp_path):
    keys = ["TTGACC", "ACGTNACGT", "ACGT", "ACGTA", "acgt", "GGGGGGGGGGGGGGGGGGGGGGGGG", "CCCC"]
    table = FakeTable(keys)
    s1 = sample([0, 2, 3], [5, 1, 7])
    s2 = sample([1, 2, 4, 5], [2, 2, 9, 4])  # key 6 ("CCCC") is counted by no sample (left by an earlier run of the table)
    df = DG.build_matrix(table, [s1, s2], ["sampleA", "sampleB"])
    assert df.index.name == "Sequence"
    assert list(df.columns) == ["annotFlag"] + DG.INITIAL_FLAGS + ["sampleA", "sampleB"]
    # rows: byte-wise lexicographic (what pandas' outer join of the per-sample frames gives), unseen keys dropped
    assert list(df.index) == sorted(k for k in keys if k != "CCCC")
    assert df.loc["ACGT", ["sampleA", "sampleB"]].tolist() == [1, 2]
    assert df.loc["acgt", ["sampleA", "sampleB"]].tolist() == [0, 9]
    assert df.loc["TTGACC", ["sampleA", "sampleB"]].tolist() == [5, 0]
    assert str(df["annotFlag"].dtype) == "int64" and (df["annotFlag"] == 0).all()
    assert all(str(df[c].dtype) == "int64" for c in ("sampleA", "sampleB"))
    assert all((df[c] == "").all() for c in DG.INITIAL_FLAGS)
    assert int(df["sampleA"].sum()) == 13 and int(df["sampleB"].sum()) == 17
    # -spl / -rr: DataFrame and the accessories tuple survive pickling unchanged
    p = tmp_path / "collapsed.pkl"
    df.to_pickle(p)
    back = pd.read_pickle(p)
    assert back.equals(df) and list(back.columns) == list(df.columns) and back.index.name == "Sequence"
    acc = ({"sampleA": 13}, {"sampleA": 13}, {"sampleA": 3}, ["/x/sampleA.fastq"], ["sampleA"])
    with open(tmp_path / "collapsed_accessories.pkl", "wb") as f:
        pickle.dump(acc, f)
    assert pickle.load(open(tmp_path / "collapsed_accessories.pkl", "rb")) == acc
    # the annotFlag split + to_csv of __main__.py:164-173 works on it
    df.loc["ACGT", "annotFlag"] = 1
    df.loc["ACGT", "exact miRNA"] = "hsa-miR-1"
    mapped, unmapped = df[df.annotFlag.eq(1)], df[df.annotFlag.eq(0)]
    mapped.to_csv(tmp_path / "mapped.csv")
    assert (tmp_path / "mapped.csv").read_text().splitlines()[0].startswith("Sequence,annotFlag,exact miRNA,")
    assert len(mapped) == 1 and len(unmapped) == 5


def test_empty_run_gives_an_empty_frame_with_the_columns():
    class Empty:
        n_keys = 0

        def export_sorted(self, seen=None):
            return np.zeros(0, dtype=np.int64), np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.uint8)

    df = DG.build_matrix(Empty(), [sample([], [])], ["s"])
    assert len(df) == 0 and list(df.columns) == ["annotFlag"] + DG.INITIAL_FLAGS + ["s"] and df.index.name == "Sequence"


def test_large_tables_get_an_arrow_index_without_python_strings():
    """Beyond ARROW_INDEX_MIN rows the 'Sequence' index is Arrow-backed (no Python object per row); it holds the same
    strings, in the same order, and writes the same CSV."""
    rng = np.random.default_rng(0)
    keys = sorted({"".join(rng.choice(list("ACGTN"), int(rng.integers(1, 40)))) for _ in range(500)})
    arr = np.array([k.encode() for k in keys], dtype="S48")
    old = DG.ARROW_INDEX_MIN
    try:
        DG.ARROW_INDEX_MIN = 10
        big = DG.sequence_index(arr)
        DG.ARROW_INDEX_MIN = 10 ** 9
        small = DG.sequence_index(arr)
    finally:
        DG.ARROW_INDEX_MIN = old
    assert small.dtype == object and "string" in str(big.dtype)
    assert big.tolist() == small.tolist() == keys and big.name == small.name == "Sequence"
    a = pd.DataFrame({"x": np.arange(len(keys))}, index=big)
    b = pd.DataFrame({"x": np.arange(len(keys))}, index=small)
    assert a.to_csv() == b.to_csv() and a.loc[keys[7], "x"] == 7
    assert (a.index.str.len().to_numpy() == np.array([len(k) for k in keys])).all()


def test_flag_columns_match_what_assign_gives():
    """empty_flag_column: the dtype and content of ``df.assign(col='')`` (digest.py:254) without n Python objects."""
    n = 1000
    ref = pd.DataFrame({"s": np.arange(n)}).assign(x="")["x"]
    col = pd.Series(DG.empty_flag_column(n), name="x")
    assert col.dtype == ref.dtype and col.equals(ref) and bool((col == "").all())


def test_device_built_annotation_column_equals_the_host_one():
    """manifoldAlign._round_column_device (Arrow offsets / data assembled with tensor ops; here on the CPU device)
    against _filled_column, incl. multi-byte names and a round without hits."""
    import torch

    from mirge_b200 import manifoldAlign as MA

    class Dev:
        tdev = torch.device("cpu")

    n = 300_000
    rng = np.random.default_rng(0)
    annot = np.full(n, 0xFF, dtype=np.uint8)
    sel = rng.random(n) < 0.2
    annot[sel] = rng.integers(0, 3, int(sel.sum()))
    annot[-1] = 1  # the last row annotated: the running sum ends inside a name
    ref = rng.integers(0, 500, n)
    names = ["hsa-miR-%d-5p/é" % i if i % 7 == 0 else "n%d" % i for i in range(500)]
    old = pd.Series(DG.empty_flag_column(n))
    for rnd in range(3):
        got = MA._round_column_device(Dev, torch.from_numpy(annot), torch.from_numpy(ref), rnd, names, old.dtype)
        rows = np.nonzero(annot == rnd)[0]
        exp = MA._filled_column(old, rows, np.asarray(names, dtype=object), ref[rows])
        if old.dtype == object:
            continue  # (pandas < 3: the device path is not used)
        assert pd.Series(got).equals(pd.Series(exp))
    assert MA._round_column_device(Dev, torch.from_numpy(annot), torch.from_numpy(ref), 5, names, old.dtype) is None
