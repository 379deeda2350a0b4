"""GPU twin of tests/test_oracle_fuzz.py: the product trim + collapse path against the C oracle on random trim
configurations, in all three kernel modes.

The file name sorts it after the other GPU tests, so a device fault here cannot take them down with it."""
import numpy as np
import pytest

from mirge_b200 import params as P
from oracle import coracle
from tests.test_oracle_fuzz import random_config, random_reads

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    from mirge_b200 import device as D

    return D.Device(0)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("seed", range(40))
def test_gpu_matches_c_oracle_on_random_configurations(dev, seed, mode):
    from mirge_b200 import device as D
    from tests.test_gpu_digest import gpu_windows, table_dict, to_dev

    rng = np.random.default_rng(9000 + seed)
    cfg = random_config(rng)
    try:
        P.build_trim_params(cfg)
    except P.UnsupportedAdapterSpec:
        pytest.skip("configuration the product rejects")
    data = random_reads(rng, cfg, 2000)
    fq = np.frombuffer(data, dtype=np.uint8)
    eng = D.DigestEngine(dev, cfg)
    eng.set_trim_mode(mode)
    n, win_o, kept_o = coracle.trim(fq, dev.trim_params)
    buf = to_dev(dev, data)
    br = eng.trim_batch(buf, buf.numel(), True)
    assert br.n_records == n
    win_g, kept_g = gpu_windows(eng, br)
    assert np.array_equal(kept_g, kept_o), (seed, cfg)
    bad = np.argwhere((win_g != win_o).any(axis=2))
    assert bad.size == 0, "seed %d %s: first differing (record, slot): %s gpu=%s oracle=%s" % (
        seed, cfg, bad[0], win_g[tuple(bad[0])], win_o[tuple(bad[0])])
    table = D.CollapseTable(dev, min_keys=256)
    eng.collapse_batch(table, br)
    _, tab = coracle.digest_collapse(fq, dev.trim_params, nthreads=2)
    if cfg.umi() is None:
        assert table_dict(table) == tab.to_dict()
