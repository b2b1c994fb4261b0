"""The parity rule of SURVEY.md section 8c, made executable.

Let s_1 >= s_2 >= ... be the oracle's scores for one query.  Position r is *decided*
iff it is separated from both neighbours by more than tau * max(|s|, 1e-6).  Decided
positions must carry exactly the oracle's id; a run of undecided positions must carry
the same ids as a set (a run that crosses the cut at k: a subset of the oracle's run);
every returned score must be within tau (relative) of the oracle's score for that id.
"""
from __future__ import annotations

import numpy as np


def check_topk_parity(prod_s, prod_i, orc_s, orc_i, k, tau, what="", floor=1e-6):
    """prod_*: [nq, k]; orc_*: [nq, >=k] (pass k + margin so the cut can be judged).

    `floor`: scores smaller in magnitude than this are compared as if they had this magnitude
    (tolerance tau * max(|s|, floor)).  bf16-vs-fp32 comparisons of cosine scores use 0.1: input
    rounding perturbs a dot product of unit vectors by an amount that does not shrink with |s|."""
    prod_s = np.asarray(prod_s, dtype=np.float64)
    prod_i = np.asarray(prod_i)
    orc_s = np.asarray(orc_s, dtype=np.float64)
    orc_i = np.asarray(orc_i)
    nq = prod_s.shape[0]
    decided = 0
    for q in range(nq):
        os_, oi = orc_s[q], orc_i[q]
        valid = int((oi >= 0).sum())
        kk = min(k, valid)
        ps, pi = prod_s[q], prod_i[q]
        assert (pi[kk:k] == -1).all(), f"{what} q={q}: expected -1 padding after {kk} hits, got {pi[kk:k]}"
        assert len(set(pi[:kk].tolist())) == kk, f"{what} q={q}: duplicate ids returned"
        score_of = {int(i): s for i, s in zip(oi[:valid], os_[:valid])}
        # runs of undecided positions
        a = 0
        while a < kk:
            b = a + 1
            while b < valid and (os_[b - 1] - os_[b]) <= tau * max(abs(os_[b - 1]), floor):
                b += 1
            run_ids = set(int(x) for x in oi[a:b])
            got = [int(x) for x in pi[a:min(b, kk)]]
            if b - a == 1:
                decided += 1
                assert got[0] == int(oi[a]), (
                    f"{what} q={q} pos={a}: id {got[0]} != oracle {int(oi[a])} "
                    f"(oracle scores {os_[max(0,a-1):a+2]}, got score {ps[a]})")
            elif b <= kk:
                assert set(got) == run_ids, f"{what} q={q} run [{a},{b}): ids {sorted(got)} != {sorted(run_ids)}"
            else:
                exhausted = (b == valid) and valid == len(oi)
                for j, g in enumerate(got):
                    if g in run_ids:
                        continue
                    assert exhausted, f"{what} q={q} run [{a},{b}) crossing k: id {g} not in oracle run"
                    # oracle list ran out inside the tie run: judge by score only
                    assert ps[a + j] >= os_[valid - 1] - tau * max(abs(os_[valid - 1]), floor), \
                        f"{what} q={q}: id {g} score {ps[a+j]} below oracle tail {os_[valid-1]}"
            a = b
        for j in range(kk):
            g = int(pi[j])
            if g in score_of:
                ref = score_of[g]
                assert abs(ps[j] - ref) <= tau * max(abs(ref), floor) + 1e-12, \
                    f"{what} q={q} pos={j} id={g}: score {ps[j]} vs oracle {ref} (tau={tau})"
        # returned list must be sorted (score desc, id asc)
        for j in range(1, kk):
            assert ps[j - 1] > ps[j] or (ps[j - 1] == ps[j] and pi[j - 1] < pi[j]), \
                f"{what} q={q}: output not sorted (score desc, id asc) at {j}: {ps[j-1]},{pi[j-1]} / {ps[j]},{pi[j]}"
    return decided
