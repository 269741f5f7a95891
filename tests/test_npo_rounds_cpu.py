"""The NPO match on the device runs in PREFIX ROUNDS (amod_b200/csrc/npo.cuh): every group of idle vehicles standing on one node
accepts, at once, the longest prefix of its preference order whose requests all name this group as their best (a request that
ties in dt with another group only as element 0).  This file restates those rounds in numpy — the same decisions the kernels
k_npo_best / k_npo_accept take, nothing of the product is imported — and checks them against the oracle's stable sort + greedy
scan (rebalancing_npo.py:26-59) on tables full of exact ties, with stations, and with more vehicles than requests and the
reverse.  It is the executable form of the argument in npo.cuh's header; the CUDA kernels themselves are checked by -m gpu."""
import numpy as np
import pytest

import oracle
from amod_b200 import synth


def npo_prefix_rounds(tt, idle, pend):
    idle, pend = np.asarray(idle), np.asarray(pend)
    nodes, inv = np.unique(idle, return_inverse=True)
    G, R = len(nodes), len(pend)
    veh = [np.flatnonzero(inv == g) for g in range(G)]           # vehicle ranks of a group, ascending
    head = np.zeros(G, int)
    D = tt[nodes[:, None] - 1, pend[None, :] - 1].astype(np.float64)
    req_match, mdt = -np.ones(R, int), np.zeros(R)
    target, matched, rounds = min(len(idle), R), 0, 0
    while matched < target:
        rounds += 1
        hrank = np.array([veh[g][head[g]] if head[g] < len(veh[g]) else -1 for g in range(G)])
        free, act = np.flatnonzero(req_match < 0), np.flatnonzero(hrank >= 0)
        best_g, tie = -np.ones(R, int), np.zeros(R, bool)
        for r in free:                                            # k_npo_best
            d = D[act, r]
            cand = act[d == d.min()]
            tie[r] = len(cand) > 1
            best_g[r] = cand[np.argmin(hrank[cand])]
        progress = 0
        for g in act:                                             # k_npo_accept
            c = len(veh[g]) - head[g]
            acc = []
            for i, (dt, r) in enumerate(sorted((D[g, r], r) for r in free)):
                if best_g[r] != g or (tie[r] and i > 0):
                    break
                acc.append((dt, r))
            for i, (dt, r) in enumerate(acc[:c]):
                req_match[r], mdt[r] = veh[g][head[g] + i], dt
            head[g] += len(acc[:c]); progress += len(acc[:c])
        assert progress > 0, "the globally smallest remaining edge is element 0 of its group: a round cannot be empty"
        matched += progress
    order = sorted(np.flatnonzero(req_match >= 0), key=lambda r: (mdt[r], r))
    return np.array([req_match[r] for r in order]), np.array(order), np.array([mdt[r] for r in order]), rounds


@pytest.mark.parametrize("kind", ["ties", "real"])
def test_prefix_rounds_equal_the_sorted_greedy_scan(kind):
    rng = np.random.default_rng(11)
    tt = synth.random_table(60, 5, kind)
    worst = 0
    for trial in range(24):
        n_idle, n_pend = int(rng.integers(1, 140)), int(rng.integers(1, 140))
        stations = rng.integers(1, 61, size=int(rng.integers(1, 8)))
        idle = np.where(rng.random(n_idle) < 0.7, rng.choice(stations, n_idle), rng.integers(1, 61, n_idle)).astype(np.int32)
        pend = rng.integers(1, 61, n_pend).astype(np.int32)
        want = oracle.npo_match(tt, idle, pend)
        got = npo_prefix_rounds(tt, idle, pend)
        for a, b in zip(want, got[:3]):
            assert np.array_equal(a, b), (kind, trial, n_idle, n_pend)
        worst = max(worst, got[3])
    assert worst < 60            # (a station with dozens of vehicles is served in a few rounds, not one vehicle per round)
