"""cutadapt >= 4's alignment objective (`compat="4"`: score = matches - mismatches - 2 x indels instead of matches;
SURVEY.md Appendix A4): the C oracle against the Python restatement, and what distinguishes the two objectives.
The restatement follows the published description of cutadapt 4; tests/test_tier3_real_tools.py holds it against an
install wherever one exists."""
import numpy as np

import mirge_b200  # noqa: F401
from mirge_b200 import params as P
from oracle import coracle
from oracle import pyoracle as po
from tests.util import ILL, py_params, random_fastq


def test_score_objective_differs_from_matches_objective_only_where_it_should():
    ad = po.Adapter("back", "ACGTACGTAC", 0.2, 3, True, True)
    rng = np.random.default_rng(0)
    differ = same = 0
    for _ in range(3000):
        read = "".join(rng.choice(list("ACGT"), int(rng.integers(8, 40))))
        if rng.random() < 0.7:  # plant a damaged copy
            p = int(rng.integers(0, len(read)))
            core = list("ACGTACGTAC")
            for _e in range(int(rng.integers(0, 3))):
                q = int(rng.integers(0, len(core)))
                r = rng.random()
                if r < 0.4:
                    core[q] = str(rng.choice(list("ACGT")))
                elif r < 0.7:
                    del core[q]
                else:
                    core.insert(q, str(rng.choice(list("ACGT"))))
            read = read[:p] + "".join(core) + read[p:]
        a, b = po.locate(ad, read, "2-3"), po.locate(ad, read, "4")
        assert (a is None) == (b is None)  # acceptance (cost within the error rate) does not depend on the objective
        if a is None:
            continue
        if a[:4] == b[:4]:
            same += 1
            # same alignment: score = matches - (mismatches + 2 indels) <= matches
            assert b[4] <= a[4] and a[5] == b[5]
        else:
            differ += 1
            # the score objective never prefers an alignment with a lower score
            assert b[4] >= a[4] - 2 * a[5] - a[5]
    assert same > 500 and differ > 0, (same, differ)


def test_c_oracle_matches_python_restatement_in_compat4():
    for name, kw in (("a", dict(adapters=[("back", ILL)])),
                     ("fb", dict(adapters=[("back", ILL), ("front", "GTTCAGAGTTCTACAGTCCGACGATC")], times=2)),
                     ("noindel", dict(adapters=[("back", ILL)], indels=False)),
                     ("wild", dict(adapters=[("back", "TGGAATTCNNGGGTGCCAAGGRACTCCAG")], error_rate=0.2, overlap=5))):
        cfg = P.TrimConfig(cutadapt_compat="4", **kw)
        data = random_fastq(700, seed=len(name) + 3, err=0.07, indel=0.05, n_rate=0.01)
        fq = np.frombuffer(data, dtype=np.uint8)
        cp = P.build_trim_params(cfg)
        assert cp.compat == 1
        n, win_c, kept_c = coracle.trim(fq, cp)
        pp = py_params(cfg)
        mods = pp.modifiers()
        lines = data.split(b"\n")
        for r in range(n):
            seq, qual = lines[4 * r + 1].decode(), lines[4 * r + 3].decode()
            wins = po.trim_windows(pp, seq, qual) if hasattr(po, "trim_windows") else None
            if wins is None:
                start, stop = 0, len(seq)
                got = []
                for mod in mods:
                    start, stop = po.apply_modifier(mod, seq, qual, start, stop, pp)
                    got.append((start, stop))
                wins = got
            E = win_c.shape[1]
            exp = wins if E == len(wins) else wins[-1:]
            for s, (a, b) in enumerate(exp):
                assert (int(win_c[r, s, 0]), int(win_c[r, s, 1])) == (a, b), (name, r, s)


def test_compat_default_and_environment(monkeypatch):
    assert P.build_trim_params(P.TrimConfig(adapters=[("back", ILL)])).compat == 0
    monkeypatch.setenv("MIRGE_B200_CUTADAPT_COMPAT", "4")
    assert P.build_trim_params(P.TrimConfig(adapters=[("back", ILL)])).compat == 1
    assert P.build_trim_params(P.TrimConfig(adapters=[("back", ILL)], cutadapt_compat="2-3")).compat == 0
