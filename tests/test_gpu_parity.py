"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and against the
reference binary's committed outputs.  Integer results must be bit-exact; output files byte-identical."""
import os
import subprocess

import numpy as np
import pytest

import fixtures
from prealpha_b200 import api, capi

pytestmark = pytest.mark.gpu
VERIFY = 0x100   # debug flag: count violations of the class-seed invariant in the fast kernel


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert api.device_count() >= 1, "GPU tests need a B200; the product has no CPU fallback"


@pytest.mark.parametrize("name", fixtures.available())
def test_fixture_histograms_and_files(name, oracle, tmp_path):
    fx = fixtures.Fixture(name)
    traj = fx.trajectory()
    jobs = fx.jobs()
    # group by input file: one multi-job call per distribution input, like the reference
    by_file = {}
    for jn, refnum, rq, cnt in jobs:
        by_file.setdefault(jn, []).append((refnum, rq, cnt))
    for jn, items in by_file.items():
        pjobs = [api.job_setup(traj, rq) for _, rq, _ in items]
        hists, within, oob, st = api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY))
        for k, (refnum, rq, cnt) in enumerate(items):
            h0, w0, o0 = oracle.histogram(traj, pjobs[k])
            assert (int(within[k]), int(oob[k])) == cnt == (w0, o0), (name, jn, refnum)
            assert np.array_equal(hists[k], h0), (name, jn, refnum)
            g, rho, rc = api.transfer(traj, pjobs[k], hists[k], within[k], oob[k])
            d = tmp_path / f"{jn}_{refnum}"
            d.mkdir()
            api.write(traj, pjobs[k], refnum, str(d) + os.sep, jn + ".inp", g, rho)
            exp = fx.expected_files(jn)
            for fn in os.listdir(d):
                assert (d / fn).read_bytes() == exp[fn], (name, jn, fn)
        assert st.kernel_launches > 0 and st.pairs_evaluated > 0


@pytest.mark.parametrize("name", ["synth", "re4_atoms"])
def test_general_kernel_equals_fast_kernel(name, oracle):
    fx = fixtures.Fixture(name)
    traj = fx.trajectory()
    pjobs = [api.job_setup(traj, rq) for _, _, rq, _ in fx.jobs() if rq.bin_b == 1 and rq.mode == capi.PA_MODE_PDF]
    pjobs = [j for j in pjobs if j.sampling_interval == pjobs[0].sampling_interval]
    a = api.histogram_multi(traj, pjobs)
    b = api.histogram_multi(traj, pjobs, api.make_exec(flags=capi.PA_EXEC_FORCE_GENERAL))
    for k in range(len(pjobs)):
        assert np.array_equal(a[0][k], b[0][k]) and a[1][k] == b[1][k] and a[2][k] == b[2][k]
    assert a[3].exception_pairs > 0 and b[3].exception_pairs == 0


def test_frame_shards_add_up(oracle):
    """Frames are independent units: shards of the sampled frames sum to the full histogram
    (what the NCCL int64 reduce relies on), for any shard count."""
    fx = fixtures.Fixture("re4")
    traj = fx.trajectory()
    pjobs = [api.job_setup(traj, rq) for _, _, rq, _ in fx.jobs()[:2]]
    full = api.histogram_multi(traj, pjobs)
    for count in (2, 3, 5, 16):
        acc = [np.zeros_like(h) for h in full[0]]
        w = np.zeros(len(pjobs), dtype=np.int64)
        for rank in range(count):
            h, wi, oo, _ = api.histogram_multi(traj, pjobs, api.make_exec(shard_rank=rank, shard_count=count))
            for k in range(len(pjobs)):
                acc[k] += h[k]
            w += wi
            for k in range(len(pjobs)):
                h0, w0, _ = oracle.histogram(traj, pjobs[k], shard=(rank, count))
                assert np.array_equal(h[k], h0) and wi[k] == w0
        for k in range(len(pjobs)):
            assert np.array_equal(acc[k], full[0][k]) and w[k] == full[1][k]


def test_tile_shards_add_up(oracle):
    """The split below the frame (pa_exec.tile_shard_*: work items of ONE frame dealt out to several GPUs) adds up to the
    oracle's histogram for every kernel family: brute-force RDF, brute-force general (2-D, cdf), charge_arm (first shard
    only), and -- combined with frame shards -- for any combination of the two splits."""
    fx = fixtures.Fixture("synth")
    traj = fx.trajectory()
    pjobs = [api.job_setup(traj, rq) for _, _, rq, _ in fx.jobs()]
    pjobs = [j for j in pjobs if j.sampling_interval == pjobs[0].sampling_interval]
    full = api.histogram_multi(traj, pjobs)
    for k, job in enumerate(pjobs):
        h0, w0, o0 = oracle.histogram(traj, job)
        assert np.array_equal(full[0][k], h0) and (int(full[1][k]), int(full[2][k])) == (w0, o0)
    for fcount, tcount in ((1, 2), (1, 5), (2, 3)):
        acc = [np.zeros_like(h) for h in full[0]]
        w = np.zeros(len(pjobs), dtype=np.int64)
        o = np.zeros(len(pjobs), dtype=np.int64)
        evals = 0
        for fr in range(fcount):
            for tr in range(tcount):
                h, wi, oo, st = api.histogram_multi(traj, pjobs, api.make_exec(shard_rank=fr, shard_count=fcount,
                                                                               tile_shard_rank=tr, tile_shard_count=tcount))
                for k in range(len(pjobs)):
                    acc[k] += h[k]
                w += wi
                o += oo
                evals += st.pairs_evaluated
        for k in range(len(pjobs)):
            assert np.array_equal(acc[k], full[0][k]) and w[k] == full[1][k] and o[k] == full[2][k], (fcount, tcount, k)
        assert evals == full[3].pairs_evaluated


def test_small_batches_and_sessions(oracle):
    fx = fixtures.Fixture("re4_atoms")
    traj = fx.trajectory()
    pjobs = [api.job_setup(traj, rq) for _, _, rq, _ in fx.jobs()[:3]]
    full = api.histogram_multi(traj, pjobs)
    tiny = api.histogram_multi(traj, pjobs, api.make_exec(frames_per_batch=1))
    for k in range(len(pjobs)):
        assert np.array_equal(full[0][k], tiny[0][k]) and full[1][k] == tiny[1][k] and full[2][k] == tiny[2][k]
    # streaming API: push frames in chunks of 5, run after each upload, download once
    s = api.Session(traj, pjobs, max_frames=5)
    for first in range(0, fx.nsteps, 5):
        s.upload(traj, first, min(5, fx.nsteps - first))
        s.run()
    h, w, o = s.download(reset=True)
    for k in range(len(pjobs)):
        assert np.array_equal(h[k], full[0][k]) and w[k] == full[1][k] and o[k] == full[2][k]
    # accumulators were reset; running the resident frames again gives just those frames
    s.run()
    h2, w2, _ = s.download()
    last = fx.nsteps - (fx.nsteps - 1) // 5 * 5
    assert int(w2[0]) > 0 and s.stats().frames_done == fx.nsteps + last
    s.close()


def _random_traj(rng, nsteps, types, L, lo=0.0, decimals=None, spread=1.5, lmp=False):
    tl = []
    for ch, na, nm, els in types:
        centre = rng.random((nsteps, nm, 1, 3)) * L * spread - L * (spread - 1) / 2 + lo
        xyz = centre + (rng.random((nsteps, nm, na, 3)) - 0.5) * 3.0
        if decimals is not None:
            xyz = np.round(xyz, decimals)
        tl.append(dict(charge=ch, natoms=na, nmol=nm, elements=els, xyz=xyz.astype(np.float32)))
    if lmp:
        box_lo = np.array([lo, lo - 0.123456789, lo + 0.5]); box_hi = box_lo + np.array([L, L * 1.1, L * 0.93])
        return capi.Trajectory(tl, box_lo, box_hi, xyz_format=False)
    lo32, hi32 = capi.cubic_box_edge(lo, lo + L)
    return capi.Trajectory(tl, lo32, hi32, xyz_format=True)


CASES = [
    # (seed, nsteps, types, L, kwargs for trajectory, requests)
    (1, 3, [(+1, 1, 700, ["Na"]), (-1, 1, 900, ["Cl"])], 40.0, dict(decimals=5),
     [dict(mode="pdf", ref_type=1, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True),
      dict(mode="pdf", ref_type=2, obs_type=2, bin_a=17, bin_b=1, maxdist=11.0),
      dict(mode="pdf", vec=(0, 0, -1), ref_type=2, obs_type=1, bin_a=400, bin_b=1, maxdist_optimize=True, weigh_charge=True)]),
    (2, 2, [(-1, 3, 300, ["O", "H", "H"]), (+2, 1, 150, ["Mg"]), (0, 5, 100, ["C", "H", "H", "H", "H"])], 25.0, dict(lmp=True),
     [dict(mode="pdf", ref_type=1, obs_type=-1, bin_a=200, bin_b=1, maxdist_optimize=True, weigh_charge=True),
      dict(mode="pdf", vec=(1, -2, 1), ref_type=3, obs_type=1, bin_a=30, bin_b=45, maxdist=9.0),
      dict(mode="cdf", vec=(0, 1, 1), ref_type=2, obs_type=-1, bin_a=40, bin_b=40, maxdist_optimize=True),
      dict(mode="pdf", ref_type=1, obs_type=1, bin_a=250, bin_b=250, maxdist=12.0)]),   # > smem histogram limit
    (3, 2, [(0, 1, 129, ["Ar"]), (0, 1, 1, ["Kr"]), (0, 1, 513, ["Xe"])], 20.0, dict(spread=4.0),
     [dict(mode="pdf", ref_type=1, obs_type=-1, bin_a=100, bin_b=1, maxdist_optimize=True),
      dict(mode="pdf", ref_type=2, obs_type=3, bin_a=100, bin_b=1, maxdist_optimize=True),
      dict(mode="pdf", ref_type=3, obs_type=2, bin_a=100, bin_b=1, maxdist_optimize=True),
      dict(mode="pdf", ref_type=2, obs_type=2, bin_a=100, bin_b=1, maxdist_optimize=True),
      dict(mode="cdf", ref_type=3, obs_type=3, bin_a=10, bin_b=10, maxdist_optimize=True)]),
]


# sparse types just above the cell-path thresholds (2000 molecules, 2e7 pairs per frame): warp tiles limited by their x extent,
# partner types below the threshold (brute force) in the same call
CASES.append((4, 2, [(+1, 1, 4300, ["Na"]), (-1, 1, 5200, ["Cl"]), (0, 1, 800, ["Ar"])], 64.0, dict(decimals=4),
              [dict(mode="pdf", ref_type=1, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True),
               dict(mode="pdf", ref_type=2, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True),
               dict(mode="pdf", ref_type=2, obs_type=1, bin_a=90, bin_b=1, maxdist=13.0),
               dict(mode="pdf", vec=(1, 0, 1), ref_type=1, obs_type=2, bin_a=30, bin_b=20, maxdist=11.0),
               dict(mode="pdf", ref_type=3, obs_type=-1, bin_a=200, bin_b=1, maxdist_optimize=True)]))


@pytest.mark.parametrize("case", CASES, ids=[f"case{c[0]}" for c in CASES])
def test_random_systems_against_oracle(case, oracle):
    seed, nsteps, types, L, tkw, reqs = case
    rng = np.random.default_rng(seed)
    traj = _random_traj(rng, nsteps, types, L, **tkw)
    pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, **r)) for r in reqs]
    hists, within, oob, st = api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY))
    for k, job in enumerate(pjobs):
        h0, w0, o0 = oracle.histogram(traj, job)
        assert (int(within[k]), int(oob[k])) == (w0, o0), (seed, k)
        assert np.array_equal(hists[k], h0), (seed, k)


def _big_system(seed, n1, n2, L, nsteps=1, clustered=False, lmp=False):
    rng = np.random.default_rng(seed)
    n = n1 + n2
    if clustered:
        centres = rng.random((12, 3)) * L
        pos = centres[rng.integers(0, 12, size=(nsteps, n))] + rng.standard_normal((nsteps, n, 3)) * (L / 14.0)
        pos[:, : n // 5] = rng.random((nsteps, n // 5, 3)) * L * 1.4 - 0.2 * L          # some uniform background, partly outside the box
    else:
        pos = rng.random((nsteps, n, 3)) * L
    pos = np.round(pos, 4)
    for s in range(nsteps):                       # planted special pairs (see tests/golden/make_golden.py::case_synth)
        pos[s, 3, :2] = pos[s, n1 + 5, :2]        # same x,y: link exactly along +-z, different types
        pos[s, 7, :2] = pos[s, 8, :2]             # same type
        pos[s, 9] = pos[s, n1 + 9]                # coincident atoms
        pos[s, 11] = (0.0, L, L / 2)              # on the faces
        pos[s, 12, 0] = pos[s, 13, 0] + np.round(L / 2, 4)   # |dx| = L/2 (image tie when L/2 is exact)
        pos[s, 12, 1:] = pos[s, 13, 1:]
    pos = pos.astype(np.float32)
    tl = [dict(charge=+1, natoms=1, nmol=n1, elements=["Na"], xyz=pos[:, :n1].reshape(nsteps, n1, 1, 3)),
          dict(charge=-1, natoms=1, nmol=n2, elements=["Cl"], xyz=pos[:, n1:].reshape(nsteps, n2, 1, 3))]
    if lmp:
        lo = np.array([-1.25, 0.333333333333, 2.0]); hi = lo + np.array([L, L * 1.07, L * 0.91])
        return capi.Trajectory(tl, lo, hi, xyz_format=False)
    lo32, hi32 = capi.cubic_box_edge(0.0, L)
    return capi.Trajectory(tl, lo32, hi32)


BIG = [("uniform", dict(seed=21, n1=13000, n2=15500, L=80.0)),
       ("clustered", dict(seed=22, n1=12600, n2=14000, L=76.0, clustered=True)),
       ("lmp_box_2frames", dict(seed=23, n1=12500, n2=12500, L=70.0, nsteps=2, lmp=True))]


@pytest.mark.parametrize("name,kw", BIG, ids=[b[0] for b in BIG])
def test_cell_sorted_kernel_against_oracle(name, kw, oracle):
    """The cell-sorted tile kernel (>= 12500 molecules per type) is bit-exact against the oracle and
    against the brute-force tile kernel, including pairs (anti)parallel to z, coincident atoms, atoms on
    the box faces and exact image ties."""
    traj = _big_system(**kw)
    reqs = [dict(mode="pdf", ref_type=1, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True),
            dict(mode="pdf", ref_type=2, obs_type=2, bin_a=300, bin_b=1, maxdist=17.25),
            dict(mode="pdf", vec=(0, 0, -1), ref_type=2, obs_type=1, bin_a=64, bin_b=1, maxdist_optimize=True, weigh_charge=True),
            # same bins as job 0 with the roles of the types swapped: its (2 around 1) pass is merged into
            # job 0's (1 around 2) pass when the arithmetic is provably exact (cubic float32 boxes)
            dict(mode="pdf", vec=(0, 0, -1), ref_type=2, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True, weigh_charge=True)]
    pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, **r)) for r in reqs]
    cells = api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY))
    nosym = api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY | capi.PA_EXEC_NO_SYMMETRY))
    brute = api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY | capi.PA_EXEC_NO_CELLS))
    assert cells[3].kernel_launches != brute[3].kernel_launches      # the cell path really ran
    if name != "lmp_box_2frames":
        assert cells[3].kernel_launches != nosym[3].kernel_launches  # ... and so did the symmetric / merged passes
    for k, job in enumerate(pjobs):
        assert np.array_equal(cells[0][k], brute[0][k]) and cells[1][k] == brute[1][k] and cells[2][k] == brute[2][k], (name, k)
        assert np.array_equal(cells[0][k], nosym[0][k]) and cells[1][k] == nosym[1][k] and cells[2][k] == nosym[2][k], (name, k)
    for k in (1, 2) if name != "lmp_box_2frames" else (1,):          # keep the CPU oracle time bounded
        h0, w0, o0 = oracle.histogram(traj, pjobs[k])
        assert np.array_equal(cells[0][k], h0) and (int(cells[1][k]), int(cells[2][k])) == (w0, o0), (name, k)
    assert int(cells[2].sum()) > 0, "the planted (anti)parallel / coincident pairs must show up as out of bounds"
    if name == "uniform":                                            # tile shards of the cell-sorted launches add up as well
        acc = [np.zeros_like(h) for h in cells[0]]
        w = np.zeros(len(pjobs), dtype=np.int64)
        o = np.zeros(len(pjobs), dtype=np.int64)
        for tr in range(3):
            h, wi, oo, _ = api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY, tile_shard_rank=tr, tile_shard_count=3))
            for k in range(len(pjobs)):
                acc[k] += h[k]
            w += wi
            o += oo
        for k in range(len(pjobs)):
            assert np.array_equal(acc[k], cells[0][k]) and w[k] == cells[1][k] and o[k] == cells[2][k], (name, k)


def test_full_size_frame_launch_combinations_agree():
    """BASELINE's full size (1M atoms in one frame, both references, 1000 bins, cutoff L/2) is far beyond the oracle's
    reach; what can be checked at that size are properties: every kernel combination (float32 launch on/off,
    symmetric + merged passes on/off) must give the same integers, the histogram must sum to `within bin`, every
    ordered pair is either binned, beyond the cutoff or out of bounds, and for uniform points the binned fraction is
    the sphere / cube volume ratio pi/6."""
    import bench
    n_half, L = bench.N_PER_TYPE, float(bench.L_BOX)
    frames = np.stack([bench.synth_frame(0, 2 * n_half, L)])
    traj = bench.make_traj(frames, n_half, L)
    jobs = bench.make_jobs(traj)
    runs = {f: api.histogram_multi(traj, jobs, api.make_exec(flags=f))
            for f in (0, capi.PA_EXEC_NO_F32, capi.PA_EXEC_NO_F32 | capi.PA_EXEC_NO_SYMMETRY)}
    ref = runs[capi.PA_EXEC_NO_F32 | capi.PA_EXEC_NO_SYMMETRY]
    for f, got in runs.items():
        for k in range(len(jobs)):
            assert np.array_equal(got[0][k], ref[0][k]) and got[1][k] == ref[1][k] and got[2][k] == ref[2][k], (f, k)
    hists, within, oob, st = runs[0]
    for k in range(len(jobs)):
        assert int(hists[k].sum()) == int(within[k])
    pairs = 2 * n_half * (2 * n_half - 1)
    assert st.pairs_evaluated == pairs
    assert abs(float(within.sum()) / pairs - np.pi / 6.0) < 1.0e-4
    assert int(oob.sum()) < 1000          # only links (anti)parallel to z within 3e-8 rad can be out of bounds


def test_cell_sorted_float32_segments_large_box(oracle):
    """A box that is large compared with a tile (L/2 = 58 A vs 40 A tiles): the tile's own cell run is a single-image,
    float32-eligible run, so the symmetric pass splits it between the fp64 launch (pairs inside the tile) and the
    float32 launch (the rest).  All launch combinations must give the same integers as the brute-force kernel, and
    the oracle agrees on the cross-type job."""
    traj = _big_system(seed=31, n1=40000, n2=12800, L=116.0)
    reqs = [dict(mode="pdf", ref_type=1, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True),
            dict(mode="pdf", vec=(0, 0, -1), ref_type=2, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True),
            dict(mode="pdf", ref_type=2, obs_type=1, bin_a=400, bin_b=1, maxdist_optimize=True)]
    pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, **r)) for r in reqs]
    runs = {f: api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY | f))
            for f in (0, capi.PA_EXEC_NO_F32, capi.PA_EXEC_NO_SYMMETRY, capi.PA_EXEC_NO_F32 | capi.PA_EXEC_NO_SYMMETRY, capi.PA_EXEC_NO_CELLS)}
    assert runs[0][3].kernel_launches > runs[capi.PA_EXEC_NO_F32][3].kernel_launches      # the float32 launches really ran
    ref = runs[capi.PA_EXEC_NO_CELLS]
    for f, got in runs.items():
        for k in range(len(pjobs)):
            assert np.array_equal(got[0][k], ref[0][k]) and got[1][k] == ref[1][k] and got[2][k] == ref[2][k], (f, k)
    h0, w0, o0 = oracle.histogram(traj, pjobs[2])
    assert np.array_equal(ref[0][2], h0) and (int(ref[1][2]), int(ref[2][2])) == (w0, o0)


@pytest.mark.parametrize("name,kw", BIG[:2] + BIG[2:], ids=[b[0] for b in BIG])
def test_cell_sorted_general_kernel(name, kw, oracle):
    """2-D pdf / cdf / centre-of-charge jobs on large systems go through the cell-sorted general kernel
    (cutoff pruning + literal evaluation of the pairs inside): identical to the brute-force general kernel
    and to the oracle."""
    traj = _big_system(**kw)
    reqs = [dict(mode="pdf", vec=(1, 1, 0), ref_type=1, obs_type=2, bin_a=24, bin_b=12, maxdist=9.0),
            dict(mode="cdf", vec=(0, 0, 1), ref_type=2, obs_type=-1, bin_a=30, bin_b=30, maxdist=11.5, weigh_charge=True),
            dict(mode="pdf", vec=(0, 0, 1), ref_type=1, obs_type=1, bin_a=300, bin_b=300, maxdist=8.0),     # 90 000 bins: global atomics
            dict(mode="pdf", vec=(0, 0, 1), ref_type=2, obs_type=1, bin_a=50, bin_b=1, maxdist=7.0, use_coc=True),
            # cutoff L/2: runs with two candidate images (fp64 segment of the float32 2-D kernel), 1000 angle bins, a skew
            # reference vector, weights
            dict(mode="pdf", vec=(2, -1, 1), ref_type=1, obs_type=-1, bin_a=40, bin_b=1000, maxdist_optimize=True, weigh_charge=True)]
    pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, **r)) for r in reqs]
    # warm-up (the first launch of a kernel includes its lazy load) with the VERIFY hook: every pair the float32 2-D kernel
    # bins without the literal evaluation is re-evaluated literally on the device and must land in the same bin
    api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY))
    cells = api.histogram_multi(traj, pjobs)
    api.histogram_multi(traj, pjobs, api.make_exec(flags=capi.PA_EXEC_NO_CELLS))
    brute = api.histogram_multi(traj, pjobs, api.make_exec(flags=capi.PA_EXEC_NO_CELLS))
    assert cells[3].kernel_launches != brute[3].kernel_launches
    for k in range(len(pjobs)):
        assert np.array_equal(cells[0][k], brute[0][k]) and cells[1][k] == brute[1][k] and cells[2][k] == brute[2][k], (name, k)
    if name == "clustered":
        for k in (0, 1, 4):
            h0, w0, o0 = oracle.histogram(traj, pjobs[k])
            assert np.array_equal(cells[0][k], h0) and (int(cells[1][k]), int(cells[2][k])) == (w0, o0), (name, k)
    if name == "uniform":
        assert cells[3].kernel_ms < brute[3].kernel_ms      # pruning must pay off at these cutoffs (homogeneous system)


def test_wrap_trajectory_switch(oracle):
    """WRAP_TRAJECTORY = T skips wrap_vector inside the distance routine (Molecular:1442)."""
    rng = np.random.default_rng(11)
    tl = [dict(charge=0, natoms=2, nmol=400, elements=["N", "N"], xyz=(rng.random((2, 400, 2, 3)) * 22.0 - 1.0).astype(np.float32))]
    lo, hi = capi.cubic_box_edge(0.0, 20.0)
    traj = capi.Trajectory(tl, lo, hi, wrap_trajectory=True)
    pjobs = [api.job_setup(traj, capi.make_request(mode=m, vec=v, ref_type=1, obs_type=1, bin_a=50, bin_b=b, maxdist_optimize=True,
                                                   sampling_interval=1)) for m, v, b in (("pdf", (0, 0, 1), 1), ("pdf", (1, 1, 1), 9), ("cdf", (0, 0, 1), 12))]
    hists, within, oob, _ = api.histogram_multi(traj, pjobs)
    for k, job in enumerate(pjobs):
        h0, w0, o0 = oracle.histogram(traj, job)
        assert np.array_equal(hists[k], h0) and (within[k], oob[k]) == (w0, o0)


def test_cutoff_off_the_last_bin_boundary(oracle):
    """maxdist = 44.0 with 1000 bins: REAL(4) step_a * 1000 is below maxdist, so the reference has pairs "inside the
    cutoff with a-bin out of range" (counted out of bounds, Distribution:829-831); maxdist = 46.5 has none.  Both take the
    approximate-class kernels, whose exact branch must tell the three outcomes apart.  A lattice with spacing chosen
    so that many pairs sit exactly in that sliver makes the case visible."""
    L = 88.0
    n_side = 25                                             # 25^3 = 15625 sites per type, spacing 3.52 A
    g = (np.arange(n_side, dtype=np.float64) + 0.25) * (L / n_side)
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(1, -1, 3)
    rng = np.random.default_rng(61)
    a = np.round(lat + rng.normal(0.0, 0.02, size=lat.shape), 4).astype(np.float32)
    # type 2: every site displaced by a vector of length ~ 44.0 (just below / above the cutoff), so that thousands of
    # pairs have d within 1e-6 relative of maxdist
    u = rng.normal(size=lat.shape); u /= np.linalg.norm(u, axis=-1, keepdims=True)
    r = 44.0 * (1.0 + rng.uniform(-4e-7, 4e-7, size=lat.shape[:2] + (1,)))
    b = (np.mod(lat + u * r, L)).astype(np.float32)
    tl = [dict(charge=+1, natoms=1, nmol=a.shape[1], elements=["Na"], xyz=a.reshape(1, -1, 1, 3)),
          dict(charge=-1, natoms=1, nmol=b.shape[1], elements=["Cl"], xyz=b.reshape(1, -1, 1, 3))]
    lo, hi = capi.cubic_box_edge(0.0, L)
    traj = capi.Trajectory(tl, lo, hi)
    reqs = [dict(mode="pdf", ref_type=1, obs_type=2, bin_a=1000, bin_b=1, maxdist_optimize=True),
            dict(mode="pdf", ref_type=2, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True)]
    pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, **r)) for r in reqs]
    assert pjobs[0].maxdist == 44.0
    runs = {f: api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY | f)) for f in (0, capi.PA_EXEC_NO_F32, capi.PA_EXEC_NO_CELLS)}
    assert runs[0][3].kernel_launches > runs[capi.PA_EXEC_NO_F32][3].kernel_launches      # approximate-class + float32 launches ran
    ref = runs[capi.PA_EXEC_NO_CELLS]
    for f, got in runs.items():
        for k in range(len(pjobs)):
            assert np.array_equal(got[0][k], ref[0][k]) and got[1][k] == ref[1][k] and got[2][k] == ref[2][k], (f, k)
    h0, w0, o0 = oracle.histogram(traj, pjobs[0])
    assert np.array_equal(ref[0][0], h0) and (int(ref[1][0]), int(ref[2][0])) == (w0, o0)
    assert o0 > 0, "the construction must produce pairs inside the cutoff whose a-bin is out of range"


@pytest.mark.parametrize("case", ["molecules", "dense_and_sparse", "orthorhombic"])
def test_float32_segments_on_inexact_coordinates(case):
    """(The cutoff 46.5 A / 1000 bins is one for which REAL(4) step_a * bin_a reaches the cutoff, i.e. the approximate-class
    variant of the cell kernel applies; e.g. 44.0 A does not, and then every pair takes the exact table.)
    Float32 segments where nothing about the coordinates is special: centres of mass of two-atom molecules (arbitrary
    doubles), types of very different density (small and large cells, tiles wider than the float32 limit), and an
    orthorhombic box with full-double bounds.  The float32 launch, the fp64-only cell kernel and the brute-force kernel
    must agree bit for bit (VERIFY re-derives every class from the fp64 table inside the kernels)."""
    rng = np.random.default_rng({"molecules": 51, "dense_and_sparse": 52, "orthorhombic": 53}[case])
    if case == "molecules":
        L, specs = 93.0, [(+1, 2, 20000), (-1, 1, 18000)]
    elif case == "dense_and_sparse":
        L, specs = 93.0, [(+1, 1, 70000), (-1, 1, 14800)]
    else:
        L, specs = 100.0, [(+1, 1, 27000), (-1, 1, 25000)]      # shortest edge 93 A -> cutoff 46.5 A as in the other cases
    lo = np.array([0.0, 0.0, 0.0]); hi = np.array([L, L, L])
    if case == "orthorhombic":
        lo = np.array([-3.3333333333, 1.25, 0.1]); hi = lo + np.array([L, L * 1.11, L * 0.93])
    tl = []
    for ch, na, nm in specs:
        c = lo + rng.random((1, nm, 1, 3)) * (hi - lo)
        xyz = c + (rng.normal(0.0, 0.45, size=(1, nm, na, 3)) if na > 1 else 0.0)
        tl.append(dict(charge=ch, natoms=na, nmol=nm, elements=["N"] * na, xyz=np.round(xyz, 5).astype(np.float32)))
    traj = capi.Trajectory(tl, lo, hi, xyz_format=(case != "orthorhombic"))
    reqs = [dict(mode="pdf", ref_type=1, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True),
            dict(mode="pdf", vec=(0, 0, -1), ref_type=2, obs_type=1, bin_a=333, bin_b=1, maxdist_optimize=True)]
    pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, **r)) for r in reqs]
    runs = {f: api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY | f)) for f in (0, capi.PA_EXEC_NO_F32, capi.PA_EXEC_NO_CELLS)}
    assert runs[0][3].kernel_launches > runs[capi.PA_EXEC_NO_F32][3].kernel_launches      # the float32 launches really ran
    ref = runs[capi.PA_EXEC_NO_CELLS]
    for f, got in runs.items():
        for k in range(len(pjobs)):
            assert np.array_equal(got[0][k], ref[0][k]) and got[1][k] == ref[1][k] and got[2][k] == ref[2][k], (case, f, k)


@pytest.mark.parametrize("wrap", [True, False])
@pytest.mark.parametrize("outside", [False, True])
def test_wrap_trajectory_with_cell_sorted_kernels(outside, wrap):
    """wrap_trajectory T on a large system: the cell-sorted kernels apply as long as every centre of mass lies inside the
    box (float32 and general kernels alike); with molecules outside the box -- slightly, absurdly far, or NaN -- the call
    must fall back to the brute-force kernels by itself, and the cell kernels that run before the flag is read must stay
    inside their tables.  Same without wrap_trajectory, where wrap_vector gives up after 1000 moves.  Both must equal
    the brute-force result (which the small-system tests pin on the oracle)."""
    rng = np.random.default_rng(41)
    n1, n2, L = 14000, 13000, 81.0
    pos = np.round(rng.random((1, n1 + n2, 3)) * L, 4)
    if outside:
        pos[0, 5] = (-0.5, 10.0, 10.0); pos[0, n1 + 7] = (40.0, L + 0.25, 3.0)
        pos[0, 6] = (1.0e7, 20.0, 20.0)                    # absurdly far (also after 1000 wrap moves): must not send a kernel out of its tables
        pos[0, n1 + 8] = (np.nan, 1.0, 1.0)
    pos = pos.astype(np.float32)
    tl = [dict(charge=+1, natoms=1, nmol=n1, elements=["Na"], xyz=pos[:, :n1].reshape(1, n1, 1, 3)),
          dict(charge=-1, natoms=1, nmol=n2, elements=["Cl"], xyz=pos[:, n1:].reshape(1, n2, 1, 3))]
    lo, hi = capi.cubic_box_edge(0.0, L)
    traj = capi.Trajectory(tl, lo, hi, wrap_trajectory=wrap)
    reqs = [dict(mode="pdf", ref_type=1, obs_type=-1, bin_a=800, bin_b=1, maxdist_optimize=True),
            dict(mode="pdf", vec=(1, 0, 1), ref_type=2, obs_type=1, bin_a=30, bin_b=18, maxdist=9.5),
            dict(mode="cdf", vec=(0, 0, 1), ref_type=1, obs_type=2, bin_a=25, bin_b=25, maxdist=8.0)]
    pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, **r)) for r in reqs]
    cells = api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY))
    brute = api.histogram_multi(traj, pjobs, api.make_exec(flags=VERIFY | capi.PA_EXEC_NO_CELLS))
    for k in range(len(pjobs)):
        assert np.array_equal(cells[0][k], brute[0][k]) and cells[1][k] == brute[1][k] and cells[2][k] == brute[2][k], k
    if not outside:
        assert cells[3].kernel_launches != brute[3].kernel_launches      # the cell path really ran
    else:
        assert wrap or int(cells[1].sum()) > 0


@pytest.mark.parametrize("name,stream", [("synth", False), ("lmp", False), ("lmp_wrap", False), ("lmp_unwrap", False), ("charge_arm", False),
                                         ("lmp_masses", False), ("synth", True), ("lmp", True), ("charge_arm", True), ("lmp_masses", True)])
def test_cli_drop_in_end_to_end(name, stream, oracle, tmp_path):
    """general.inp -> molecular.inp -> xyz / lmp text -> GPU -> output files: byte-identical to what the
    reference binary wrote for the same inputs (tests/golden), same `within bin` / `out of bounds` lines."""
    fx = fixtures.Fixture(name)
    wd = tmp_path / "run"
    (wd / "out").mkdir(parents=True)
    # %.9g round-trips every float32 exactly, so the CLI parses the very coordinates the binary parsed
    if fx.lmp and fx.general_extra:
        # lmp_masses: `masses` / `atomic_masses` sections in molecular.inp, three extra columns per atom read with
        # `extra_velocity T`, trajectory type taken from the file extension
        with open(wd / "traj.lmp", "w") as f:
            n = len(fx.elements)
            for s_ in range(fx.coords.shape[0]):
                f.write(f"ITEM: TIMESTEP\n{s_ * 1000}\nITEM: NUMBER OF ATOMS\n{n}\nITEM: BOX BOUNDS pp pp pp\n")
                for k in range(3):
                    f.write(f"{float(fx.box_lo[k])!r} {float(fx.box_hi[k])!r}\n")
                f.write("ITEM: ATOMS element xu yu zu vx vy vz\n")
                f.write("".join(f"{fx.elements[a]} {fx.coords[s_, a, 0]:.9g} {fx.coords[s_, a, 1]:.9g} {fx.coords[s_, a, 2]:.9g} 1.5e-3 -2.0 7\n" for a in range(n)))
        tname, box_line = "traj.lmp", ""
    elif fx.lmp:
        oracle.write_lmp(str(wd / "traj.lmp"), fx.elements, fx.coords.astype(np.float64), fx.box_lo, fx.box_hi, fmt="%.9g")
        tname, box_line = "traj.lmp", ""
    else:
        oracle.write_xyz(str(wd / "traj.xyz"), fx.elements, fx.coords.astype(np.float64), fmt="%.9g")
        tname, box_line = "traj.xyz", f" cubic_box_edge {float(fx.box_lo[0])!r} {float(fx.box_hi[0])!r}\n"
    charges = ""
    if fx.atom_charges is not None:
        charges = f"atomic_charges {sum(len(q) for q in fx.atom_charges)}\n" + "".join(
            f" {ti + 1} {ai + 1} {float(v)!r}\n" for ti, q in enumerate(fx.atom_charges) for ai, v in enumerate(q))
    (wd / "molecular.inp").write_text(f" {fx.nsteps}\n {len(fx.types)}\n" + "".join(f" {c:+d} {a} {m}\n" for c, a, m in fx.types) + charges + fx.mol_tail + "quit\n")
    switch = " wrap_trajectory T\n" if fx.wrap else (" unwrap_trajectory T\n" if fx.unwrap else "")
    if fx.general_extra:
        switch += f" {fx.general_extra}\n"
    gen = f' "{tname}"\n "molecular.inp"\n "./"\n "./"\n "./out/"\n{box_line}{switch} sequential_read {"T" if stream else "F"}\n set_threads 8\n'
    for jn, text in zip(fx.job_names, fx.job_texts):
        (wd / f"{jn}.inp").write_text(text)
        gen += f' set_prefix "{jn}_"\n distribution "{jn}.inp"\n'
    gen += " gyradius -1 1 1\n quit\n"
    (wd / "general.inp").write_text(gen)
    cli = os.path.join(os.path.dirname(capi.LIB_PATH), "prealpha_b200_cli")
    env = dict(os.environ)
    if stream:
        env["PA_STREAM_BATCH"] = "3"          # sequential_read T: stream the file three frames at a time
    p = subprocess.run([cli, "general.inp"], cwd=wd, capture_output=True, timeout=600, env=env)
    out = p.stdout.decode()
    assert p.returncode == 0, out[-3000:] + p.stderr.decode()[-2000:]
    assert ("streams the trajectory from the file" in out) == stream
    counts = oracle.parse_within(out)
    assert counts == fx.counts
    produced = {fn: (wd / "out" / fn).read_bytes() for fn in os.listdir(wd / "out")}
    assert set(produced) == set(fx.files)
    for fn, content in produced.items():
        assert content == fx.files[fn], fn
    if name in ("synth", "lmp", "charge_arm"):
        # binary frame cache: the first run with PREALPHA_B200_CACHE writes it, the second reads frames from it instead of
        # parsing text (also through the streaming path); with two GPUs or more the frames / batches are dealt out to all
        # of them.  Same bytes in every output file.
        env["PREALPHA_B200_CACHE"] = "1"
        runs = [([cli, "general.inp"], False), ([cli, "general.inp"], True)]
        if api.device_count() >= 2:
            runs.append(([cli, "-g", "all", "general.inp"], True))
        for cmd, cached in runs:
            for fn in os.listdir(wd / "out"):
                os.remove(wd / "out" / fn)
            p = subprocess.run(cmd, cwd=wd, capture_output=True, timeout=600, env=env)
            out = p.stdout.decode()
            assert p.returncode == 0, out[-3000:] + p.stderr.decode()[-2000:]
            assert ("reading the binary frame cache" in out) == cached and ("wrote the binary frame cache" in out) == (not cached)
            assert oracle.parse_within(out) == fx.counts
            again = {fn: (wd / "out" / fn).read_bytes() for fn in os.listdir(wd / "out")}
            assert again == produced, cmd


def test_exception_list_overflow_falls_back(oracle):
    """A perfect lattice has whole columns of atoms sharing (x,y): every pair of a column is (anti)parallel to
    z.  With a tiny exception list (test hook 0x200) the fast kernels overflow and the call is redone with
    the general kernel; the result must not change."""
    n = 12
    g = np.arange(n) * 2.0 + 0.5
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(3)
    pts = pts[rng.permutation(len(pts))]
    half = len(pts) // 2
    tl = [dict(charge=+1, natoms=1, nmol=half, elements=["Na"], xyz=pts[None, :half].reshape(1, half, 1, 3)),
          dict(charge=-1, natoms=1, nmol=half, elements=["Cl"], xyz=pts[None, half:].reshape(1, half, 1, 3))]
    lo, hi = capi.cubic_box_edge(0.0, 24.0)
    traj = capi.Trajectory(tl, lo, hi)
    job = api.job_setup(traj, capi.make_request(mode="pdf", ref_type=1, obs_type=-1, bin_a=120, bin_b=1, maxdist_optimize=True, sampling_interval=1))
    h0, w0, o0 = oracle.histogram(traj, job)
    assert o0 > 2048
    normal = api.histogram_multi(traj, [job])
    tiny = api.histogram_multi(traj, [job], api.make_exec(flags=0x200))
    for res in (normal, tiny):
        assert np.array_equal(res[0][0], h0) and (int(res[1][0]), int(res[2][0])) == (w0, o0)
    assert normal[3].exception_pairs > 0 and tiny[3].exception_pairs == 0      # general kernel: no exception list


def test_degenerate_inputs(oracle):
    rng = np.random.default_rng(8)
    xyz1 = (rng.random((3, 1, 1, 3)) * 10).astype(np.float32)
    xyz2 = (rng.random((3, 40, 2, 3)) * 10).astype(np.float32)
    xyz2[1, 7, 0, 1] = np.nan                                 # a NaN coordinate: never inside the cutoff, nothing crashes
    tl = [dict(charge=0, natoms=1, nmol=1, elements=["Ar"], xyz=xyz1), dict(charge=0, natoms=2, nmol=40, elements=["N", "N"], xyz=xyz2)]
    lo, hi = capi.cubic_box_edge(0.0, 10.0)
    traj = capi.Trajectory(tl, lo, hi)
    reqs = [dict(mode="pdf", ref_type=1, obs_type=1, bin_a=1, bin_b=1, maxdist_optimize=True),          # one molecule: no pairs at all
            dict(mode="pdf", ref_type=1, obs_type=2, bin_a=1000, bin_b=1, maxdist=4.99),
            dict(mode="pdf", ref_type=2, obs_type=-1, bin_a=1, bin_b=1000, maxdist=1e-3),               # cutoff so small that nothing is binned
            dict(mode="cdf", ref_type=2, obs_type=2, bin_a=1000, bin_b=1, maxdist_optimize=True)]
    for si in (1, 2, 7):                                      # sampling interval larger than the trajectory: frame 1 only
        pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=si, **r)) for r in reqs]
        hists, within, oob, st = api.histogram_multi(traj, pjobs)
        assert st.frames_done == (3 + si - 1) // si
        assert within[0] == 0 and oob[0] == 0 and not hists[0].any()
        for k, job in enumerate(pjobs):
            h0, w0, o0 = oracle.histogram(traj, job)
            assert np.array_equal(hists[k], h0) and (int(within[k]), int(oob[k])) == (w0, o0), (si, k)


def test_error_codes():
    import ctypes as C
    fx = fixtures.Fixture("synth")
    traj = fx.trajectory()
    lib = capi.load()
    job = api.job_setup(traj, fx.jobs()[0][2])
    hist = np.zeros(job.bin_a * job.bin_b, dtype=np.int64)
    w, o = C.c_int64(0), C.c_int64(0)
    bad = capi.pa_job(); C.memmove(C.byref(bad), C.byref(job), C.sizeof(capi.pa_job)); bad.ref_type = 9
    assert lib.pa_distribution_histogram(traj.ptr, C.byref(bad), hist.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(w), C.byref(o)) == 33
    bad.ref_type = 1; bad.bin_a = 0
    assert lib.pa_distribution_histogram(traj.ptr, C.byref(bad), hist.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(w), C.byref(o)) == capi.PA_ERR_BAD_ARG
    rq = capi.make_request(vec=(0, 0, 0), sampling_interval=1)          # randomised reference vector: undefined in the reference
    j2 = capi.pa_job()
    assert lib.pa_job_setup(traj.ptr, C.byref(rq), C.byref(j2)) == capi.PA_ERR_UNSUPPORTED
    rq = capi.make_request(ref_type=3, sampling_interval=1)
    assert lib.pa_job_setup(traj.ptr, C.byref(rq), C.byref(j2)) == 33


def test_cli_rejects_short_trajectories(oracle, tmp_path):
    """perform_distribution_analysis refuses trajectories with fewer than 11 steps (error 37, Distribution:1191)."""
    wd = tmp_path
    rng = np.random.default_rng(0)
    oracle.write_xyz(str(wd / "t.xyz"), ["Ar"] * 20, rng.random((5, 20, 3)) * 9.0)
    (wd / "molecular.inp").write_text(" 5\n 1\n +0 1 20\nquit\n")
    (wd / "d.inp").write_text("pdf 1\n+0 +0 +1 1 1\nmaxdist_optimize\nquit\n")
    (wd / "general.inp").write_text(' "t.xyz"\n "molecular.inp"\n "./"\n "./"\n "./"\n cubic_box_edge 0.0 9.0\n distribution "d.inp"\n nonsense_keyword\n quit\n')
    cli = os.path.join(os.path.dirname(capi.LIB_PATH), "prealpha_b200_cli")
    p = subprocess.run([cli, "general.inp"], cwd=wd, capture_output=True, timeout=120)
    out = p.stdout.decode()
    assert p.returncode == 0 and "ERROR 37" in out and "ERROR 20" in out          # exit status 0 like the binary's: errors go to the report
    assert not [f for f in os.listdir(wd) if f.startswith("reference_")]


@pytest.mark.parametrize("no_nccl", [False, True], ids=["nccl_reduce", "peer_copy_reduce"])
def test_multi_device_call_matches_single_device(no_nccl, tmp_path):
    """pa_exec.ndevices: one call drives several GPUs (one host thread each) and combines the int64 accumulators on the
    first device with ncclReduce (or peer copies + add when NCCL is switched off).  Frames >= devices: frame shards;
    a single frame: tile shards.  Same integers as one device either way.  Needs >= 2 GPUs (skipped on a 1-GPU box);
    run in a subprocess because the NCCL switch is read once per process."""
    if api.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import fixtures
from prealpha_b200 import api, capi
from test_gpu_parity import _big_system
nd = api.device_count()
fx = fixtures.Fixture("re4")
traj = fx.trajectory()
pjobs = [api.job_setup(traj, rq) for _, _, rq, _ in fx.jobs()[:2]]
one = api.histogram_multi(traj, pjobs)
for n in sorted({2, nd, -1}):
    many = api.histogram_multi(traj, pjobs, api.make_exec(ndevices=n))
    for k in range(len(pjobs)):
        assert np.array_equal(one[0][k], many[0][k]) and one[1][k] == many[1][k] and one[2][k] == many[2][k], (n, k)
    assert many[3].frames_done == one[3].frames_done and many[3].pairs_evaluated == one[3].pairs_evaluated
# one frame: the devices split the work items of the frame (cell-sorted kernels)
traj = _big_system(seed=21, n1=13000, n2=15500, L=80.0)
pjobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, mode="pdf", ref_type=1, obs_type=-1, bin_a=1000, bin_b=1, maxdist_optimize=True)),
         api.job_setup(traj, capi.make_request(sampling_interval=1, mode="pdf", vec=(1, 1, 0), ref_type=2, obs_type=1, bin_a=24, bin_b=12, maxdist=9.0))]
one = api.histogram_multi(traj, pjobs)
many = api.histogram_multi(traj, pjobs, api.make_exec(ndevices=nd))
for k in range(len(pjobs)):
    assert np.array_equal(one[0][k], many[0][k]) and one[1][k] == many[1][k] and one[2][k] == many[2][k], k
assert many[3].frames_done == 1 and many[3].pairs_evaluated == one[3].pairs_evaluated
print("ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    if no_nccl:
        env["PA_NO_NCCL"] = "1"
    else:
        env["PA_NCCL_MIN_BYTES"] = "0"                 # small accumulators default to peer copies: force the ncclReduce path
    out = subprocess.run([os.sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


def test_large_batches_equal_single_frame_batches():
    """Batches of tens of MB are copied into the page-locked staging buffer by several host threads (pa_session_upload):
    40 frames of 70 000 atoms in batches of 20 frames (16.8 MB each, two copy threads, exactness scan included) against
    the same frames one per batch; a frame with a coordinate outside the exactness range sits in the middle of a batch."""
    traj = _big_system(seed=33, n1=34000, n2=36000, L=110.0, nsteps=40)
    traj.xyz[1][23, 17, 0, 2] = np.float32(3.0e-6)               # |z| < 2^-17: this batch must take the non-symmetric passes
    jobs = [api.job_setup(traj, capi.make_request(sampling_interval=1, mode="pdf", ref_type=r, obs_type=-1, bin_a=400, bin_b=1, maxdist_optimize=True))
            for r in (1, 2)]
    big = api.histogram_multi(traj, jobs, api.make_exec(frames_per_batch=20))
    one = api.histogram_multi(traj, jobs, api.make_exec(frames_per_batch=1))
    assert big[3].frames_done == one[3].frames_done == 40 and big[3].kernel_launches < one[3].kernel_launches
    for k in range(2):
        assert np.array_equal(big[0][k], one[0][k]) and big[1][k] == one[1][k] and big[2][k] == one[2][k], k
