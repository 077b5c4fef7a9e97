"""the neighbour test of the cluster / speciation analyses (REAL(4) minimum-image atom distance against
REAL(4) squared cutoffs).  CPU: the oracle's literal loops against an independent numpy evaluation of the same 27-image
minimum; GPU: the CUDA primitive against the oracle, pair list identical and in the same order."""
import numpy as np
import pytest

import fixtures
from prealpha_b200 import api, capi


def _instances(fx, entries):
    """(type, atom) entries -> all atom instances, entry-major, molecule inner (as the reference's lists are built)."""
    out, cls = [], []
    for e, (ty, at) in enumerate(entries):
        nm = fx.types[ty - 1][2]
        out += [(ty, m, at) for m in range(1, nm + 1)]
        cls += [e] * nm
    return np.array(out, dtype=np.int32), np.array(cls, dtype=np.int32)


def _numpy_pairs(fx, traj_box, step, a, b, cut, a_cls, b_cls, same_list, skip_same):
    """Independent restatement: 27 images in the reference's order on wrapped fp64 positions, float32 result."""
    lo, hi = traj_box
    L = hi - lo
    off = np.cumsum([0] + [na * nm for _, na, nm in fx.types])

    def pos(lst):
        idx = [off[t - 1] + (m - 1) * fx.types[t - 1][1] + (at - 1) for t, m, at in lst]
        p = fx.coords[step - 1][idx].astype(np.float64)
        for _ in range(1000):
            low, high = p < lo, p > hi
            if not (low.any() or high.any()):
                break
            p = np.where(low, p + L, np.where(high, p - L, p))
        return p
    pa_, pb = pos(a), pos(b)
    mds = float(np.sum(L * L))
    pairs = []
    for ia in range(len(a)):
        best = np.full(len(b), mds)
        for sx in (-1.0, 0.0, 1.0):
            for sy in (-1.0, 0.0, 1.0):
                for sz in (-1.0, 0.0, 1.0):
                    d = (pb + np.array([sx, sy, sz]) * L) - pa_[ia]
                    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
                    best = np.where(d2 < best, d2, best)
        hit = best.astype(np.float32) < cut[a_cls[ia], b_cls]
        for ib in np.nonzero(hit)[0]:
            if same_list and ib <= ia:
                continue
            if skip_same and a[ia][0] == b[ib][0] and a[ia][1] == b[ib][1]:
                continue
            pairs.append((ia, int(ib)))
    return np.array(pairs, dtype=np.int32).reshape(-1, 2)


def _cases(fx):
    a, a_cls = _instances(fx, [(1, 1), (1, 2), (2, 1)])                 # O, H of water-like type 1; Na
    b, b_cls = _instances(fx, [(4, 1), (3, 1), (1, 3)])                 # C of type 4; Mg; the other H
    cut = (np.array([[3.1, 2.4, 1.6], [2.2, 3.3, 1.1], [4.0, 2.9, 3.6]], dtype=np.float32) ** 2).astype(np.float32)
    yield "speciation-like", dict(a=a, b=b, cutoff_sq=cut, a_class=a_cls, b_class=b_cls, skip_same_molecule=True)
    allat, _ = _instances(fx, [(1, 1), (2, 1), (3, 1), (4, 2)])
    yield "cluster-like", dict(a=allat, b=None, cutoff_sq=np.float32(2.75) ** 2, same_list=True)


def test_neighbour_oracle_matches_independent_numpy(oracle):
    fx = fixtures.Fixture("lmp")
    traj = fx.trajectory()
    box = (np.asarray(fx.box_lo, dtype=np.float64), np.asarray(fx.box_hi, dtype=np.float64))
    for name, kw in _cases(fx):
        rq, keep = capi.make_neigh_request(3, **kw)
        got = oracle.neighbour_pairs(traj, rq)
        a = kw["a"]; b = a if kw.get("same_list") else kw["b"]
        cut = np.atleast_2d(np.asarray(kw["cutoff_sq"], dtype=np.float32))
        a_cls = kw.get("a_class", np.zeros(len(a), dtype=np.int32)); b_cls = kw.get("b_class", np.zeros(len(b), dtype=np.int32))
        exp = _numpy_pairs(fx, box, 3, a, b, cut, a_cls, b_cls, kw.get("same_list", False), kw.get("skip_same_molecule", False))
        assert len(exp) > 20, name
        assert np.array_equal(got, exp), name


@pytest.mark.gpu
def test_neighbour_pairs_cuda_matches_oracle(oracle):
    fx = fixtures.Fixture("lmp")
    traj = fx.trajectory()
    for step in (1, 7):
        for name, kw in _cases(fx):
            rq, keep = capi.make_neigh_request(step, **kw)
            exp = oracle.neighbour_pairs(traj, rq)
            got, tests = api.neighbour_pairs(traj, rq, capacity=16)          # forces the "call again with a larger buffer" path
            assert np.array_equal(got, exp), (name, step)
            assert tests > 0
    # errors of the reference: unknown molecule type (33), molecule (110) and atom (111) indices
    for bad, code in (((9, 1, 1), 33), ((1, 10 ** 6, 1), 110), ((1, 1, 99), 111)):
        rq, keep = capi.make_neigh_request(1, a=np.array([bad], dtype=np.int32), b=np.array([(1, 1, 1)], dtype=np.int32), cutoff_sq=np.float32(1.0))
        with pytest.raises(RuntimeError):
            api.neighbour_pairs(traj, rq)


@pytest.mark.gpu
def test_neighbour_pairs_larger_random_system(oracle):
    rng = np.random.default_rng(5)
    types = []
    for ch, na, nm in [(1, 1, 2500), (-1, 2, 1500)]:
        c = rng.random((1, nm, 1, 3)) * 44.0 - 2.0
        types.append(dict(charge=ch, natoms=na, nmol=nm, elements=["C"] * na,
                          xyz=np.round(c + rng.normal(0, 0.6, size=(1, nm, na, 3)), 4).astype(np.float32)))
    traj = capi.Trajectory(types, [0.0] * 3, [40.0, 41.5, 39.25])
    inst = np.array([(1, m, 1) for m in range(1, 2501)] + [(2, m, a) for m in range(1, 1501) for a in (1, 2)], dtype=np.int32)
    rq, keep = capi.make_neigh_request(1, a=inst, b=None, cutoff_sq=np.float32(3.0) ** 2, same_list=True, skip_same_molecule=True)
    exp = oracle.neighbour_pairs(traj, rq)
    got, tests = api.neighbour_pairs(traj, rq)                                    # cell list: 13 x 13 x 13 cells of >= 3 A
    brute, tests_brute = api.neighbour_pairs(traj, rq, exec_=api.make_exec(flags=0x800))
    assert len(exp) > 1000 and np.array_equal(got, exp) and np.array_equal(brute, exp)
    assert tests_brute == len(inst) * (len(inst) - 1) // 2 and 0 < tests < tests_brute // 20


@pytest.mark.gpu
def test_neighbour_cell_list_equals_brute_force_on_100k_atoms():
    """The size SURVEY 8f-4 names (all-atom list of a 100k-atom box against itself, 3 A cutoff): the cell-list kernel
    returns exactly the brute-force kernel's pairs (both apply the same literal REAL(4) predicate), with per-class
    cutoffs, positions on the box faces and outside the box (wrapped by wrap_vector first), and it does ~1e-3 of the
    distance tests."""
    rng = np.random.default_rng(7)
    n, L = 50_000, 126.0
    pos = np.round(rng.random((2, n, 3)) * L, 4)
    pos[0, :50] = pos[1, :50] + 0.5                      # close pairs across the two types
    pos[0, 50:60, 0] = 0.0                               # on the lower face
    pos[1, 50:60, 0] = L                                 # on the upper face (periodic neighbours of the ones above)
    pos[1, 60:70, 1] += L                                # outside the box: wrapped before the search
    types = [dict(charge=+1, natoms=1, nmol=n, elements=["Na"], xyz=pos[0].reshape(1, n, 1, 3).astype(np.float32)),
             dict(charge=-1, natoms=1, nmol=n, elements=["Cl"], xyz=pos[1].reshape(1, n, 1, 3).astype(np.float32))]
    lo, hi = capi.cubic_box_edge(0.0, L)
    traj = capi.Trajectory(types, lo, hi)
    inst = np.array([(t, m, 1) for t in (1, 2) for m in range(1, n + 1)], dtype=np.int32)
    cls = np.array([0] * n + [1] * n, dtype=np.int32)
    cut = (np.array([[3.0, 2.6], [2.6, 3.2]], dtype=np.float32) ** 2).astype(np.float32)
    rq, keep = capi.make_neigh_request(1, a=inst, b=None, cutoff_sq=cut, a_class=cls, b_class=cls, same_list=True)
    got, tests = api.neighbour_pairs(traj, rq)
    brute, tests_brute = api.neighbour_pairs(traj, rq, exec_=api.make_exec(flags=0x800))
    assert len(brute) > 100_000 and np.array_equal(got, brute)
    assert tests < tests_brute // 500
