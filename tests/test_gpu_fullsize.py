"""Parity at BASELINE.json's full sizes (1M-atom RDF frame, 4M-atom 2-D histogram).

The oracle evaluates 27 images per pair at ~6e6 pairs/s per host core, so a full frame (1e12 ordered pairs) is out of
its reach.  What is pinned here, on exactly the frames the benchmark uses:
  * the oracle's own loop nest (OpenMP over blocks of reference molecules, `oracle_lib.histogram_frame`) against the
    CUDA path for reference types cut out of the frame: 12 500 molecules (cell-sorted kernels) and 1 900 molecules
    (brute-force tile kernels) around ALL 1M atoms -- same box, same cutoff, same 1000 bins as the headline line;
  * the cell-sorted float32 / symmetric / merged launches of the whole frame against the brute-force kernels of the
    whole frame (the kernels pinned on the oracle by the rows above and by every smaller test), bit for bit, with the
    VERIFY hook re-deriving every class from the fp64 table;
  * tile shards (pa_exec.tile_shard_*: the i-range split of one frame over several GPUs) adding up to the whole.
Marked `big`: `-m "gpu and not big"` skips them while iterating (they need a few minutes of host CPU)."""
import numpy as np
import pytest

import bench
from prealpha_b200 import api, capi

pytestmark = [pytest.mark.gpu, pytest.mark.big]
VERIFY = 0x100


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert api.device_count() >= 1, "GPU tests need a B200; the product has no CPU fallback"


def _cut_traj(frame, cuts, L):
    """The benchmark frame [natoms][3] as one-atom types of the given sizes (in file order)."""
    tl, off = [], 0
    for k, n in enumerate(cuts):
        tl.append(dict(charge=+1 if k % 2 == 0 else -1, natoms=1, nmol=n, elements=["Na" if k % 2 == 0 else "Cl"],
                       xyz=frame[off:off + n].reshape(1, n, 1, 3)))
        off += n
    assert off == frame.shape[0]
    lo, hi = capi.cubic_box_edge(0.0, L)
    return capi.Trajectory(tl, lo, hi, xyz_format=True)


def _rdf_job(traj, ref, obs):
    return api.job_setup(traj, capi.make_request(mode="pdf", vec=(0, 0, 1), ref_type=ref, obs_type=obs, bin_a=bench.BINS, bin_b=1,
                                                 maxdist_optimize=True, sampling_interval=1))


@pytest.mark.parametrize("nref,expect_cells", [(12500, True), (1900, False)], ids=["cells_12500_rows", "brute_1900_rows"])
def test_headline_frame_reference_rows_against_oracle(nref, expect_cells, oracle):
    """`nref` reference atoms of the 1M-atom benchmark frame around all 1M atoms: histogram, within and out-of-bounds
    counts equal the oracle's (1.25e10 resp. 1.9e9 literal 27-image pair evaluations on the host cores)."""
    n_half, L = bench.N_PER_TYPE, float(bench.L_BOX)
    frame = bench.synth_frame(0, 2 * n_half, L)
    traj = _cut_traj(frame, [nref, n_half - nref, n_half], L)
    job = _rdf_job(traj, 1, -1)
    h, w, o, st = api.histogram_multi(traj, [job], api.make_exec(flags=VERIFY))
    assert (st.cell_pair_evals > 0) == expect_cells            # which kernel family did the work
    h0, w0, o0 = oracle.histogram_frame(traj, job, step=0)
    assert (int(w[0]), int(o[0])) == (w0, o0)
    assert np.array_equal(h[0], h0)
    assert st.pairs_evaluated == nref * (2 * n_half - 1)


def test_headline_frame_cells_equal_brute_force_and_tile_shards_add_up():
    """The whole benchmark frame, both reference lines: cell-sorted float32 + symmetric + merged launches (what bench.py
    times) == brute-force tile kernels, and 3 tile shards (i-range split of ONE frame) sum to the same integers."""
    n_half, L = bench.N_PER_TYPE, float(bench.L_BOX)
    frames = np.stack([bench.synth_frame(0, 2 * n_half, L)])
    traj = bench.make_traj(frames, n_half, L)
    jobs = bench.make_jobs(traj)
    cells = api.histogram_multi(traj, jobs, api.make_exec(flags=VERIFY))
    brute = api.histogram_multi(traj, jobs, api.make_exec(flags=capi.PA_EXEC_NO_CELLS))
    assert cells[3].cell_pair_evals > 0 and brute[3].cell_pair_evals == 0
    for k in range(len(jobs)):
        assert np.array_equal(cells[0][k], brute[0][k]) and cells[1][k] == brute[1][k] and cells[2][k] == brute[2][k], k
    # the kernels' own counters: every counted pair was evaluated once (symmetric / merged passes), 2 ordered pairs each
    # (pairs whose link is (anti)parallel to z go through the exception kernel and are not in the kernels' counter)
    assert 0 <= int(cells[1].sum()) - 2 * cells[3].unique_pairs_counted <= cells[3].exception_pairs
    assert cells[3].unique_pairs_counted <= cells[3].smem_atomics <= cells[3].cell_pair_evals < n_half * (2 * n_half)
    acc = [np.zeros_like(h) for h in cells[0]]
    w = np.zeros(len(jobs), dtype=np.int64)
    o = np.zeros(len(jobs), dtype=np.int64)
    evals = 0
    for rank in range(3):
        hh, ww, oo, st = api.histogram_multi(traj, jobs, api.make_exec(tile_shard_rank=rank, tile_shard_count=3))
        for k in range(len(jobs)):
            acc[k] += hh[k]
        w += ww
        o += oo
        evals += st.cell_pair_evals
        assert 0.25 < st.cell_pair_evals / cells[3].cell_pair_evals < 0.42, "tile shards should be balanced"
    for k in range(len(jobs)):
        assert np.array_equal(acc[k], cells[0][k]) and w[k] == cells[1][k] and o[k] == cells[2][k], k
    assert evals == cells[3].cell_pair_evals


C5_L = 430.887        # 4M atoms at 0.05 A^-3


def _c5_frame():
    g = np.random.Generator(np.random.Philox(key=[12345, 5_000_000]))
    return np.round(g.random((4_000_000, 3)) * C5_L, 5).astype(np.float32)


def _c5_job(traj, ref, obs):
    return api.job_setup(traj, capi.make_request(mode="pdf", vec=(0, 0, 1), ref_type=ref, obs_type=obs, bin_a=200, bin_b=90,
                                                 maxdist=40.0, sampling_interval=1))


def test_c5_frame_reference_rows_against_oracle(oracle):
    """BASELINE config 5 (4M atoms, 200x90 distance-angle histogram, maxdist 40 A): 1 000 reference atoms of the frame
    around all 4M atoms against the oracle (4e9 literal pair evaluations), through the brute-force general kernel."""
    frame = _c5_frame()
    nref = 1000
    traj = _cut_traj(frame, [nref, 2_000_000 - nref, 2_000_000], C5_L)
    job = _c5_job(traj, 1, -1)
    h, w, o, st = api.histogram_multi(traj, [job])
    h0, w0, o0 = oracle.histogram_frame(traj, job, step=0)
    assert (int(w[0]), int(o[0])) == (w0, o0)
    assert np.array_equal(h[0], h0)
    assert int(w[0]) > 10_000_000


def test_c5_frame_cells_equal_brute_force_and_tile_shards_add_up():
    """The whole 4M-atom frame, reference type 1 (2M atoms) around all: cell-sorted general kernel == brute-force
    general kernel, and 4 tile shards add up to the same 18 000-bin histogram."""
    frame = _c5_frame()
    traj = _cut_traj(frame, [2_000_000, 2_000_000], C5_L)
    job = _c5_job(traj, 1, -1)
    api.histogram_multi(traj, [job], api.make_exec(flags=VERIFY))        # every float32-binned pair re-evaluated literally on the device
    cells = api.histogram_multi(traj, [job])
    brute = api.histogram_multi(traj, [job], api.make_exec(flags=capi.PA_EXEC_NO_CELLS))
    assert cells[3].kernel_ms < brute[3].kernel_ms / 5
    assert np.array_equal(cells[0][0], brute[0][0]) and cells[1][0] == brute[1][0] and cells[2][0] == brute[2][0]
    acc, w, o = np.zeros_like(cells[0][0]), 0, 0
    for rank in range(4):
        hh, ww, oo, _ = api.histogram_multi(traj, [job], api.make_exec(tile_shard_rank=rank, tile_shard_count=4))
        acc += hh[0]
        w += int(ww[0])
        o += int(oo[0])
    assert np.array_equal(acc, cells[0][0]) and (w, o) == (int(cells[1][0]), int(cells[2][0]))
