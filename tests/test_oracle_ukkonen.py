"""cutadapt's ``Aligner.locate`` does not fill whole DP columns: it keeps ``last``, the index of the lowest cell whose cost
is at most k = int(rate * m), computes a column only down to there, lets ``last`` grow by one row per column and looks at
row m only while ``last == m`` (Ukkonen's cut-off).  Cells below ``last`` keep whatever an earlier column left in them.
The oracle (oracle/pyoracle.py::locate) and the kernels fill the full column and claim that no accepted result depends on
the difference.  This file holds that claim to a literal restatement of the control flow -- stale
cells included -- on the adversarial inputs of tests/test_adapter_search_host.py, for every placement, with and without
indels and wildcards.  (Restated from cutadapt 2.x-3.x ``_align.pyx`` like the oracle itself: it pins the equivalence of
the two forms, not the third-party semantics.)"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.test_adapter_search_host import WHERE_SPECS, make_adapter, make_read, placed_read


def locate_with_cutoff(ad: po.Adapter, read: str):
    m, n = len(ad.sequence), len(read)
    rate = ad.max_error_rate
    ins_cost = del_cost = 1 if ad.indels else po.INDEL_OFF_COST
    up = read.upper()
    table = po.IUPAC if ad.read_wildcards else (po.ACGT if ad.wildcard_ref else po.ACGT_ASCII)
    s2 = [table.get(ch, 0) for ch in up]
    s1 = ad.masks
    start_in_ref, stop_in_ref, start_in_query, stop_in_query = po.WHERE_FLAGS[ad.where]
    k = int(rate * m)
    max_n = n if start_in_query else min(n, m + k)
    min_n = 0 if stop_in_query else max(0, n - m - k)
    cost, origin, matches = [0] * (m + 1), [0] * (m + 1), [0] * (m + 1)
    for i in range(m + 1):
        if not start_in_ref and not start_in_query:
            cost[i], origin[i] = max(i, min_n) * ins_cost, 0
        elif start_in_ref and not start_in_query:
            cost[i], origin[i] = min_n * ins_cost, min(0, min_n - i)
        elif not start_in_ref and start_in_query:
            cost[i], origin[i] = i * ins_cost, max(0, min_n - i)
        else:
            cost[i], origin[i] = min(i, min_n) * ins_cost, min_n - i
    best = dict(cost=m + n, origin=0, matches=0, ref_stop=m, query_stop=n)
    last = m if start_in_ref else min(m, k + 1)  # Ukkonen: the last cell that can be at most k
    stopped = False

    def eff(length, i):
        if not ad.wildcard_ref:
            return length
        if length < m:
            ref_start = -min(origin[i], 0)
            return length - (ad.n_counts[i] - ad.n_counts[ref_start])
        return ad.effective_length

    for j in range(min_n + 1, max_n + 1):
        dc, do, dm = cost[0], origin[0], matches[0]
        if start_in_query:
            origin[0] = j
        else:
            cost[0] = j * ins_cost
        for i in range(1, last + 1):
            if s1[i - 1] & s2[j - 1]:
                c, o, mt = dc, do, dm + 1
            else:
                cd, cdel, cins = dc + 1, cost[i] + del_cost, cost[i - 1] + ins_cost
                if cd <= cdel and cd <= cins:
                    c, o, mt = cd, do, dm
                elif cins <= cdel:
                    c, o, mt = cins, origin[i - 1], matches[i - 1]
                else:
                    c, o, mt = cdel, origin[i], matches[i]
            dc, do, dm = cost[i], origin[i], matches[i]
            cost[i], origin[i], matches[i] = c, o, mt
        while last >= 0 and cost[last] > k:
            last -= 1
        if last < m:
            last += 1
        elif stop_in_query:
            length = m + min(origin[m], 0)
            if length < m and ad.wildcard_ref:
                e = length - (ad.n_counts[m] - ad.n_counts[m - length])
            else:
                e = ad.effective_length if ad.wildcard_ref else length
            c, mt = cost[m], matches[m]
            if length >= ad.min_overlap and c <= e * rate and (mt > best["matches"] or (mt == best["matches"] and c < best["cost"])):
                best.update(matches=mt, cost=c, origin=origin[m], ref_stop=m, query_stop=j)
                if c == 0 and mt == m:
                    stopped = True
                    break
    if not stopped and max_n == n:
        for i in range(0 if stop_in_ref else m, m + 1):
            length = i + min(origin[i], 0)
            c, mt = cost[i], matches[i]
            if length >= ad.min_overlap and c <= eff(length, i) * rate and (mt > best["matches"] or (mt == best["matches"] and c < best["cost"])):
                best.update(matches=mt, cost=c, origin=origin[i], ref_stop=i, query_stop=n)
    if best["cost"] == m + n:
        return None
    o = best["origin"]
    return (0 if o >= 0 else -o, best["ref_stop"], o if o >= 0 else 0, best["query_stop"], best["matches"], best["cost"])


@pytest.mark.parametrize("seed", range(10))
def test_full_columns_equal_the_cut_off_form(seed):
    rng = np.random.default_rng(5200 + seed)
    wheres = sorted(WHERE_SPECS)
    n_found = n_total = 0
    for rep in range(18):
        where = wheres[(seed + rep) % len(wheres)]
        adseq = make_adapter(rng, ["random", "homopolymer", "repeat", "two_blocks", "prefix_repeat"][rep % 5])
        if rng.random() < 0.3:
            adseq = list(adseq)
            adseq[int(rng.integers(len(adseq)))] = "N"
            adseq = "".join(adseq)
            if set(adseq) == {"N"}:
                adseq = "C" + adseq
        ad = po.Adapter(where, adseq, float(rng.choice([0.0, 0.05, 0.1, 0.12, 0.2, 0.34])), int(rng.integers(1, 8)),
                        bool(rng.random() < 0.7), True, bool(rng.random() < 0.3))
        plain = "".join(c if c in "ACGT" else "C" for c in adseq)
        reads = [placed_read(rng, plain, where) for _ in range(25)] + [make_read(rng, plain, int(rng.integers(0, 6))) for _ in range(25)]
        for read in reads:
            if not read:
                continue
            exp = po.locate(ad, read)
            got = locate_with_cutoff(ad, read)
            assert got == exp, (where, adseq, ad.max_error_rate, ad.min_overlap, ad.indels, ad.read_wildcards, read, got, exp)
            n_total += 1
            n_found += exp is not None
    assert n_total > 700 and n_found > 150, (n_total, n_found)
