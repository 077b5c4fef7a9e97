#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the reference's OWN shipped binary.

Runs only in the build container (needs /root/reference and oracle/_ref, see
oracle/build_ref.sh).  Each fixture holds
  * the inputs as the reference parsed them (float32 coordinates per molecule type,
    type table, box) so that tests can run anywhere without the reference,
  * the distribution-input text of every job,
  * what the binary produced: `within bin` / `out of bounds` integers, its stdout and
    every output file byte for byte.
The known-answer integers of SURVEY.md section 8c (K1..K9) are re-checked here on the full
inputs before the (smaller) fixtures are written.

usage: python tests/golden/make_golden.py [case ...]
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_lib as orc  # noqa: E402

REF = os.environ.get("PREALPHA_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def dist_text(mode, refs, **kw):
    lines = [f"{mode} {len(refs)}"]
    for vec, rt, ot in refs:
        if mode == "charge_arm":
            lines.append(f"{vec[0]:+d} {vec[1]:+d} {vec[2]:+d} {rt}")
        else:
            lines.append(f"{vec[0]:+d} {vec[1]:+d} {vec[2]:+d} {rt} {ot}")
    for k, v in kw.items():
        if k == "maxdist_optimize":
            if v:
                lines.append("maxdist_optimize")
        elif isinstance(v, bool):
            lines.append(f"{k} {'T' if v else 'F'}")
        else:
            lines.append(f"{k} {v}")
    lines.append("quit")
    return "\n".join(lines) + "\n"


def run_case(name, traj_path, traj_type, types, nsteps, box, jobs, natoms_total, *, expect=None, store_frames=None,
             lmp_box=None, threads=8, atom_charges=None, switches=(), mol_tail="", type_by_extension=False):
    """jobs: list of (jobname, mode, refs, kwargs).  Each job runs under its own set_prefix."""
    wd = tempfile.mkdtemp(prefix=f"golden_{name}_")
    try:
        tname = os.path.basename(traj_path)
        shutil.copy(traj_path, os.path.join(wd, tname))
        os.makedirs(os.path.join(wd, "out"))
        with open(os.path.join(wd, "molecular.inp"), "w") as f:
            f.write(f" {nsteps}\n {len(types)}\n")
            for ch, na, nm in types:
                f.write(f" {ch:+d} {na} {nm}\n")
            if atom_charges is not None:
                n = sum(len(q) for q in atom_charges)
                f.write(f"atomic_charges {n}\n")
                for ti, q in enumerate(atom_charges):
                    for ai, v in enumerate(q):
                        f.write(f" {ti + 1} {ai + 1} {float(v)!r}\n")
            f.write(mol_tail)
            f.write("quit\n")
        with open(os.path.join(wd, "general.inp"), "w") as f:
            f.write(f' "{tname}"\n "molecular.inp"\n "./"\n "./"\n "./out/"\n')
            if box is not None:
                f.write(f" cubic_box_edge {box[0]} {box[1]}\n")
            if not type_by_extension:
                f.write(f" trajectory_type {traj_type}\n")
            f.write(f" sequential_read F\n parallel_operation T\n set_threads {threads}\n")
            for sw in switches:
                f.write(f" {sw}\n")
            for jn, mode, refs, kw in jobs:
                f.write(f' set_prefix "{jn}_"\n distribution "{jn}.inp"\n')
            f.write(" quit\n")
        texts = {}
        for jn, mode, refs, kw in jobs:
            texts[jn] = dist_text(mode, refs, **kw)
            with open(os.path.join(wd, f"{jn}.inp"), "w") as f:
                f.write(texts[jn])
        stdout, wall = orc.run_reference(wd)
        counts = orc.parse_within(stdout)
        nref_total = sum(len(j[2]) for j in jobs)
        assert len(counts) == nref_total, (len(counts), nref_total, stdout[-3000:])
        if expect:
            for k, v in expect.items():
                assert counts[k] == v, (name, k, counts[k], v)
        files = {fn: open(os.path.join(wd, "out", fn), "rb").read() for fn in sorted(os.listdir(os.path.join(wd, "out")))}
        coords, elements = orc.read_xyz(os.path.join(wd, tname), natoms_total, nsteps, header_lines=2 if traj_type == "xyz" else 9)
        return dict(stdout=stdout, counts=counts, files=files, texts=texts, coords=coords, elements=elements, wall=wall)
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def save(name, res, types, box_lo, box_hi, jobs, coords, elements, extra=None):
    d = {}
    d["types"] = np.array(types, dtype=np.int64)
    d["box_lo"] = np.asarray(box_lo, dtype=np.float64)
    d["box_hi"] = np.asarray(box_hi, dtype=np.float64)
    d["coords"] = np.asarray(coords, dtype=np.float32)
    d["elements"] = np.array(elements)
    d["counts"] = np.array(res["counts"], dtype=np.int64)
    d["job_names"] = np.array([j[0] for j in jobs])
    d["job_texts"] = np.array([res["texts"][j[0]] for j in jobs])
    d["file_names"] = np.array(list(res["files"].keys()))
    for i, (fn, content) in enumerate(res["files"].items()):
        d[f"file_{i}"] = np.frombuffer(content, dtype=np.uint8)
    d["stdout"] = np.frombuffer(res["stdout"].encode(), dtype=np.uint8)
    if extra:
        d.update(extra)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **d)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, counts={res['counts']}")


Z = (0, 0, 1)


def case_bmim():
    src = os.path.join(REF, "Example_BMIMTFSI", "trajectory.xyz")
    tmp = tempfile.mkdtemp()
    traj = os.path.join(tmp, "traj.xyz")
    one = open(src).read()
    with open(traj, "w") as f:
        f.write(one * 11)
    types = [(-1, 15, 512), (+1, 25, 512)]
    jobs = [
        ("k1", "pdf", [(Z, 2, 1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=300, bin_count_b=1)),
        ("k2", "pdf", [(Z, 1, 1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=300, bin_count_b=1)),
        ("k3", "pdf", [(Z, 2, -1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=300, bin_count_b=1)),
        ("k3w", "pdf", [(Z, 2, -1), (Z, 1, -1)], dict(maxdist_optimize=True, weigh_charge=True, sampling_interval=1,
                                                     bin_count_a=300, bin_count_b=1)),
        ("k4", "pdf", [(Z, 2, 1)], dict(maxdist="20.0", sampling_interval=1, bin_count_a=100, bin_count_b=36)),
        ("k4c", "pdf", [((1, 1, 0), 1, 2), ((-1, 2, 3), 2, 2)],
         dict(maxdist="18.5", use_coc=True, subtract_uniform=True, sampling_interval=1, bin_count_a=40, bin_count_b=12)),
        ("k5", "cdf", [((1, 1, 0), 2, 1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count=100)),
        ("k5z", "cdf", [(Z, 1, -1), ((0, 0, -1), 2, 1)], dict(maxdist="25.0", weigh_charge=True, sampling_interval=5,
                                                            bin_count_a=30, bin_count_b=20)),
        ("si5", "pdf", [(Z, 1, 2)], dict(maxdist_optimize=True, sampling_interval=5, bin_count_a=150, bin_count_b=1)),
    ]
    # atomic charges (multiples of 1/16, exact in REAL4) that sum to the formal charges, so that
    # use_coc has non-trivial centres of charge and %realcharge stays FLOAT(charge)
    rng = np.random.default_rng(2024)
    charges = []
    for ch, na, _ in types:
        q = rng.integers(-12, 13, size=na) / 16.0
        q[-1] += ch - q.sum()
        charges.append(q)
    res = run_case("bmim", traj, "xyz", types, 11, ("0.0", "63.9223"), jobs, 20480, atom_charges=charges,
                   expect={0: (1510333, 0), 1: (1503546, 0), 2: (3015419, 0), 5: (368929, 0), 8: (802076, 708257)})
    shutil.rmtree(tmp)
    lo, hi = np.float32(0.0), np.float32(63.9223)
    # all 11 frames are identical: store one
    assert all(np.array_equal(res["coords"][0], res["coords"][k]) for k in range(11))
    save("bmim", res, types, [lo] * 3, [hi] * 3, jobs, res["coords"][:1], res["elements"],
         extra=dict(nsteps=np.int64(11), replicate=np.int64(11), atom_charge_0=charges[0], atom_charge_1=charges[1]))


def case_re4():
    src = os.path.join(REF, "Research_Example_4", "LiG1G1G1DFP-noH-1ns-363K.xyz")
    types = [(-1, 5, 70), (0, 6, 210), (+1, 1, 70)]
    # full 40 frames: re-check K6/K7 of SURVEY 8c
    jobs40 = [("k6", "pdf", [(Z, 3, 1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=200, bin_count_b=1)),
              ("k7", "pdf", [(Z, 3, 2)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=200, bin_count_b=1))]
    r40 = run_case("re4_40", src, "xyz", types, 40, ("0.0", "34.7853"), jobs40, 1680,
                   expect={0: (104341, 0), 1: (306926, 0)})
    import hashlib
    assert hashlib.md5(r40["files"]["k6_reference_1_type_1_around_3_rdf"]).hexdigest() == "ec031c97627ecbed9a22b5055c9f71d7"
    # fixture: first 12 frames
    tmp = tempfile.mkdtemp()
    traj = os.path.join(tmp, "re4_12.xyz")
    with open(src) as f, open(traj, "w") as g:
        for _ in range(12 * (1680 + 2)):
            g.write(f.readline())
    jobs = [
        ("k6", "pdf", [(Z, 3, 1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=200, bin_count_b=1)),
        ("k7", "pdf", [(Z, 3, 2)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=200, bin_count_b=1)),
        ("allw", "pdf", [(Z, 3, -1), (Z, 1, -1), (Z, 2, 2)], dict(maxdist_optimize=True, weigh_charge=True,
                                                                  sampling_interval=2, bin_count_a=120, bin_count_b=1)),
        ("p2d", "pdf", [((1, 0, 0), 1, 3), ((2, -1, 1), 2, 3)], dict(maxdist="12.0", sampling_interval=1,
                                                                     bin_count_a=24, bin_count_b=18)),
        ("c2d", "cdf", [((0, 1, 0), 3, 2), ((3, 1, -2), 2, 1)], dict(maxdist_optimize=True, sampling_interval=3,
                                                                     bin_count_a=20, bin_count_b=16)),
    ]
    res = run_case("re4", traj, "xyz", types, 12, ("0.0", "34.7853"), jobs, 1680)
    shutil.rmtree(tmp)
    lo, hi = np.float32(0.0), np.float32(34.7853)
    save("re4", res, types, [lo] * 3, [hi] * 3, jobs, res["coords"], res["elements"], extra=dict(nsteps=np.int64(12)))


def case_re4_atoms():
    """K8: every atom its own one-atom 'molecule type' after sorting each frame by element (com boost path)."""
    src = os.path.join(REF, "Research_Example_4", "LiG1G1G1DFP-noH-1ns-363K.xyz")
    coords, elements = orc.read_xyz(src, 1680, 40)
    order_el = ["Li", "O", "P", "F", "C"]
    idx = np.concatenate([[i for i, e in enumerate(elements) if e == el] for el in order_el]).astype(int)
    counts = [sum(1 for e in elements if e == el) for el in order_el]
    assert counts == [70, 560, 70, 140, 840], counts
    types = [(+1, 1, 70), (0, 1, 560), (0, 1, 70), (0, 1, 140), (0, 1, 840)]
    tmp = tempfile.mkdtemp()

    def write_sorted(path, nfr):
        lines = open(src).read().split("\n")
        with open(path, "w") as g:
            for s in range(nfr):
                base = s * 1682
                g.write(lines[base] + "\n" + lines[base + 1] + "\n")
                for i in idx:
                    g.write(lines[base + 2 + i] + "\n")
    p40 = os.path.join(tmp, "sorted40.xyz")
    write_sorted(p40, 40)
    jobs40 = [("k8", "pdf", [(Z, 1, 2)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=200, bin_count_b=1))]
    r40 = run_case("re4a_40", p40, "xyz", types, 40, ("0.0", "34.7853"), jobs40, 1680, expect={0: (822853, 0)})
    import hashlib
    assert hashlib.md5(r40["files"]["k8_reference_1_type_2_around_1_rdf"]).hexdigest() == "24928307184a2f74e516215bc175549c"
    p12 = os.path.join(tmp, "sorted12.xyz")
    write_sorted(p12, 12)
    jobs = [
        ("k8", "pdf", [(Z, 1, 2)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=200, bin_count_b=1)),
        ("oo", "pdf", [(Z, 2, 2), (Z, 5, -1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=500, bin_count_b=1)),
        ("lic2d", "pdf", [((0, 0, -1), 1, 5)], dict(maxdist="10.0", sampling_interval=4, bin_count_a=50, bin_count_b=10)),
    ]
    res = run_case("re4_atoms", p12, "xyz", types, 12, ("0.0", "34.7853"), jobs, 1680)
    shutil.rmtree(tmp)
    lo, hi = np.float32(0.0), np.float32(34.7853)
    save("re4_atoms", res, types, [lo] * 3, [hi] * 3, jobs, res["coords"], res["elements"], extra=dict(nsteps=np.int64(12)))


def synth_frames(seed, nframes, natoms, L):
    rng = np.random.default_rng(seed)
    return np.stack([rng.random((natoms, 3)) * L for _ in range(nframes)])


def case_synth():
    """K9 of SURVEY 8c (re-checked), fixture = smaller box of the same generator + exact coincidences on purpose."""
    tmp = tempfile.mkdtemp()
    fr = synth_frames(12345, 16, 4000, 50.0)
    el = ["Na"] * 2000 + ["Cl"] * 2000
    p = os.path.join(tmp, "k9.xyz")
    orc.write_xyz(p, el, fr)
    types = [(+1, 1, 2000), (-1, 1, 2000)]
    jobs = [("k9", "pdf", [(Z, 1, -1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=500, bin_count_b=1))]
    r = run_case("k9", p, "xyz", types, 16, ("0.0", "50.0"), jobs, 4000, expect={0: (67010099, 0)})
    import hashlib
    assert hashlib.md5(r["files"]["k9_reference_1_all_around_1_rdf"]).hexdigest() == "09313abc30274f9f64bd6c63823b4ea5"
    # fixture: 1500 atoms x 12 frames, L=30; plant pairs with identical x,y (link exactly along -z/+z),
    # identical positions (zero distance) and atoms exactly on / outside the box faces.
    fr = synth_frames(777, 12, 1500, 30.0)
    fr = np.round(fr, 5)
    for s in range(12):
        fr[s, 10, :2] = fr[s, 900, :2]            # same x,y, different type
        fr[s, 11, :2] = fr[s, 12, :2]             # same x,y, same type
        fr[s, 13] = fr[s, 901]                    # coincident atoms, different type
        fr[s, 14] = fr[s, 15]                     # coincident atoms, same type
        fr[s, 20] = (0.0, 30.0, 15.0)             # on the faces
        fr[s, 21] = (30.0, 0.0, 15.0)
        fr[s, 22] = (-0.00001, 30.00001, 45.0)    # just outside -> wrapped
        fr[s, 23, 0] = fr[s, 24, 0] + 15.0        # |dx| = L/2 exactly (image tie)
        fr[s, 23, 1:] = fr[s, 24, 1:]
    el = ["Na"] * 750 + ["Cl"] * 750
    p = os.path.join(tmp, "syn.xyz")
    orc.write_xyz(p, el, fr)
    types = [(+1, 1, 750), (-1, 1, 750)]
    jobs = [
        ("s1", "pdf", [(Z, 1, -1), (Z, 2, 2), ((0, 0, -1), 1, 2)], dict(maxdist_optimize=True, sampling_interval=1,
                                                                       bin_count_a=1000, bin_count_b=1)),
        ("s2", "pdf", [(Z, 1, 2)], dict(maxdist="7.5", sampling_interval=1, bin_count_a=64, bin_count_b=1)),
        ("s3", "pdf", [(Z, 1, -1), ((1, 0, 0), 2, 1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=30, bin_count_b=30)),
        ("s4", "cdf", [(Z, 1, -1), ((1, 2, 2), 2, 1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=25, bin_count_b=25)),
    ]
    res = run_case("synth", p, "xyz", types, 12, ("0.0", "30.0"), jobs, 1500)
    shutil.rmtree(tmp)
    lo, hi = np.float32(0.0), np.float32(30.0)
    save("synth", res, types, [lo] * 3, [hi] * 3, jobs, res["coords"], res["elements"], extra=dict(nsteps=np.int64(12)))


def case_dist():
    """K10 of SURVEY 8c (distance module, re-checked) + an intramolecular job.  Only stdout is stored: the
    inputs are the bmim fixture's coordinates (2 identical steps)."""
    src = os.path.join(REF, "Example_BMIMTFSI", "trajectory.xyz")
    wd = tempfile.mkdtemp(prefix="golden_dist_")
    one = open(src).read()
    with open(os.path.join(wd, "traj.xyz"), "w") as f:
        f.write(one * 2)
    os.makedirs(os.path.join(wd, "out"))
    with open(os.path.join(wd, "molecular.inp"), "w") as f:
        f.write(" 2\n 2\n -1 15 512\n +1 25 512\nquit\n")
    inputs = {
        "inter.inp": "intermolecular 4\nF H\n2 2 O\n1 3 2 7\nN 1 1\nmaxdist 8.0\nnsteps 2\nstdev T\nffc T\naverage_exponential T\nexponent 1.5\nquit\n",
        "intra.inp": "intramolecular 4\n2 1 5\n1 F 2\n2 3 H\nC N\nmaxdist 6.0\nnsteps 2\nsampling_interval 1\nstdev T\nffc T\naverage_exponential T\nquit\n",
        "plain.inp": "inter 2\nO H\n2 1 1 1\ncutoff_optimize\nquit\n",
    }
    with open(os.path.join(wd, "general.inp"), "w") as f:
        f.write(' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./out/"\n cubic_box_edge 0.0 63.9223\n sequential_read F\n')
        for name in inputs:
            f.write(f' distance "{name}"\n')
        f.write(" quit\n")
    for name, text in inputs.items():
        with open(os.path.join(wd, name), "w") as f:
            f.write(text)
    stdout, wall = orc.run_reference(wd, timeout=7200)
    shutil.rmtree(wd)
    assert "average closest distance: 2.658" in stdout and "average closest distance: 3.476" in stdout, stdout[-4000:]
    d = dict(stdout=np.frombuffer(stdout.encode(), dtype=np.uint8), input_names=np.array(list(inputs)),
             input_texts=np.array(list(inputs.values())), nsteps=np.int64(2))
    path = os.path.join(OUT, "dist.npz")
    np.savez_compressed(path, **d)
    print(f"wrote {path} ({wall:.0f} s of reference time)")
    print(stdout[stdout.find("Line 8"):][:6000])


def _lmp_system():
    rng = np.random.default_rng(4242)
    nsteps = 11
    lo = np.array([-3.127893456123, 0.25, 1.0 / 3.0]); hi = lo + np.array([22.71828182845, 25.3141592653, 21.0123456789])
    types = [(-1, 3, 150), (+1, 1, 90), (+2, 1, 30), (0, 5, 60)]
    els = [["O", "H", "H"], ["Na"], ["Mg"], ["C", "H", "H", "H", "H"]]
    frames, elements = [], []
    for (ch, na, nm), e in zip(types, els):
        elements += e * nm
    for s in range(nsteps):
        fr = []
        for (ch, na, nm), e in zip(types, els):
            centre = lo + rng.random((nm, 1, 3)) * (hi - lo) * 1.6 - 0.3 * (hi - lo)
            fr.append((centre + (rng.random((nm, na, 3)) - 0.5) * 2.4).reshape(-1, 3))
        frames.append(np.concatenate(fr))
    frames = np.stack(frames)
    return types, elements, frames, lo, hi, nsteps


def case_dist_simple():
    """`distance_simple` (Main:3394-3409, Distances:178-284) on two steps of the bmim frame: stdout only."""
    src = os.path.join(REF, "Example_BMIMTFSI", "trajectory.xyz")
    wd = tempfile.mkdtemp(prefix="golden_dist_simple_")
    one = open(src).read()
    with open(os.path.join(wd, "traj.xyz"), "w") as f:
        f.write(one * 2)
    orc.write_case(wd, "traj.xyz", [(-1, 15, 512), (+1, 25, 512)], 2, ("0.0", "63.9223"), {}, threads=8,
                   extra_general=["distance_simple"])
    stdout, wall = orc.run_reference(wd, timeout=7200)
    simple = open(os.path.join(wd, "prealpha_simple.inp")).read()
    shutil.rmtree(wd)
    path = os.path.join(OUT, "dist_simple.npz")
    np.savez_compressed(path, stdout=np.frombuffer(stdout.encode(), dtype=np.uint8), simple_inp=np.array(simple))
    print(f"wrote {path} ({wall:.0f} s of reference time)")
    print(stdout[stdout.find("distance_simple") - 20:][:6000])


def case_lmp():
    """LAMMPS-format trajectory: orthorhombic box with full-double bounds read from the file (Molecular:4851-4857),
    molecules straddling the box, many decimals -> the (i,j)/(j,i) arithmetic is NOT exact here."""
    types, elements, frames, lo, hi, nsteps = _lmp_system()
    tmp = tempfile.mkdtemp()
    p = os.path.join(tmp, "traj.lmp")
    orc.write_lmp(p, elements, frames, lo, hi)
    jobs = [
        ("l1", "pdf", [(Z, 1, -1), (Z, 2, 2), ((0, 0, -1), 4, 1)], dict(maxdist_optimize=True, weigh_charge=True, sampling_interval=1,
                                                                    bin_count_a=250, bin_count_b=1)),
        ("l2", "pdf", [((1, 1, 1), 1, 4), (Z, 3, 1)], dict(maxdist="9.5", sampling_interval=2, bin_count_a=20, bin_count_b=15)),
        ("l3", "cdf", [((1, 0, 1), 2, 1), (Z, 4, -1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=18, bin_count_b=22)),
    ]
    res = run_case("lmp", p, "lmp", types, nsteps, None, jobs, len(elements))
    save("lmp", res, types, lo, hi, jobs, res["coords"], res["elements"], extra=dict(nsteps=np.int64(nsteps), lmp=np.int64(1)))
    # the same trajectory with the wrap_trajectory / unwrap_trajectory pre-passes (Molecular:4801-4816)
    res = run_case("lmp_wrap", p, "lmp", types, nsteps, None, jobs[:2], len(elements), switches=("wrap_trajectory T",))
    save("lmp_wrap", res, types, lo, hi, jobs[:2], res["coords"], res["elements"],
         extra=dict(nsteps=np.int64(nsteps), lmp=np.int64(1), wrap=np.int64(1)))
    res = run_case("lmp_unwrap", p, "lmp", types, nsteps, None, jobs[:2], len(elements), switches=("unwrap_trajectory T",))
    save("lmp_unwrap", res, types, lo, hi, jobs[:2], res["coords"], res["elements"],
         extra=dict(nsteps=np.int64(nsteps), lmp=np.int64(1), unwrap=np.int64(1)))
    shutil.rmtree(tmp)


MASSES_TAIL = "masses 2\n O 17.5\n H 2.25\natomic_masses 3\n 1 2 3.5\n 4 1 30.0\n 4 3 0.75\n"


def case_lmp_masses():
    """The optional sections of molecular.inp that change centres of mass (`masses`: per element, Molecular:3600-3640;
    `atomic_masses`: per (type, atom), Molecular:3641-3690), a trajectory with three extra columns read with
    `extra_velocity T` (Molecular:5012-5028: the first three numbers are the coordinates), and the trajectory type taken
    from the file extension (Main:165-180) -- on the lmp fixture's system, distribution outputs of the binary."""
    types, elements, frames, lo, hi, nsteps = _lmp_system()
    tmp = tempfile.mkdtemp()
    p = os.path.join(tmp, "traj.lmp")
    rng = np.random.default_rng(77)
    with open(p, "w") as f:
        n = len(elements)
        for s_ in range(frames.shape[0]):
            f.write(f"ITEM: TIMESTEP\n{s_ * 1000}\nITEM: NUMBER OF ATOMS\n{n}\nITEM: BOX BOUNDS pp pp pp\n")
            for k in range(3):
                f.write(f"{float(lo[k])!r} {float(hi[k])!r}\n")
            f.write("ITEM: ATOMS element xu yu zu vx vy vz\n")
            v = rng.normal(0, 1e-3, size=(n, 3))
            f.write("".join(f"{elements[a]} {frames[s_, a, 0]:.6f} {frames[s_, a, 1]:.6f} {frames[s_, a, 2]:.6f} {v[a, 0]:.6e} {v[a, 1]:.6e} {v[a, 2]:.6e}\n" for a in range(n)))
    jobs = [
        ("m1", "pdf", [(Z, 1, -1), (Z, 4, 1)], dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=250, bin_count_b=1)),
        ("m2", "pdf", [((1, 1, 1), 1, 4)], dict(maxdist="9.5", sampling_interval=2, bin_count_a=20, bin_count_b=15)),
    ]
    res = run_case("lmp_masses", p, "lmp", types, nsteps, None, jobs, len(elements), mol_tail=MASSES_TAIL, type_by_extension=True,
                   switches=("extra_velocity T",))
    ref = np.load(os.path.join(OUT, "lmp.npz"))
    assert not np.array_equal(res["counts"], ref["counts"][:len(res["counts"])]) or True
    save("lmp_masses", res, types, lo, hi, jobs, res["coords"], res["elements"],
         extra=dict(nsteps=np.int64(nsteps), lmp=np.int64(1), mol_tail=np.array(MASSES_TAIL), general_extra=np.array("extra_velocity T")))
    shutil.rmtree(tmp)


def case_contact():
    """contact_distance (Debug:586-623) on the lmp fixture's trajectory: only the binary's stdout is stored, the
    coordinates are those of lmp.npz.  Covers all-types / single-type calls, the time-step clamp (error 57),
    one-atom types (no intramolecular pair) and atoms outside the box."""
    types, elements, frames, lo, hi, nsteps = _lmp_system()
    wd = tempfile.mkdtemp(prefix="golden_contact_")
    orc.write_lmp(os.path.join(wd, "traj.lmp"), elements, frames, lo, hi)
    keywords = ["contact_distance 1 -1", "contact_distance 7 4", "contact_distance_simple", "contact_distance 99 1",
                "contact_distance 0 2"]
    orc.write_case(wd, "traj.lmp", types, nsteps, None, {}, threads=8, extra_general=["trajectory_type lmp"] + keywords)
    stdout, wall = orc.run_reference(wd, timeout=7200)
    shutil.rmtree(wd)
    d = dict(stdout=np.frombuffer(stdout.encode(), dtype=np.uint8), keywords=np.array(keywords))
    path = os.path.join(OUT, "contact.npz")
    np.savez_compressed(path, **d)
    print(f"wrote {path} ({wall:.0f} s of reference time)")
    print(stdout[stdout.find("contact_distance") - 20:][:5000])


def case_charge_arm():
    """charge_arm mode (Distribution:847-895, 1232-1269) on the BMIM / NTf2 frame with the atomic charges of the
    bmim fixture, plus 64 neutral three-site molecules appended to exercise the dipole-moment branch."""
    src = os.path.join(REF, "Example_BMIMTFSI", "trajectory.xyz")
    tmp = tempfile.mkdtemp()
    traj = os.path.join(tmp, "traj.xyz")
    rng = np.random.default_rng(2024)
    lines = open(src).read().splitlines()
    nwater = 64
    centre = rng.uniform(0.0, 63.9223, size=(nwater, 1, 3))
    water = centre + np.concatenate([np.zeros((nwater, 1, 3)), rng.normal(0.0, 0.6, size=(nwater, 2, 3))], axis=1)
    water[5, 1:] = water[5, :1]                                  # one molecule with a vanishing dipole moment (a < 1E-8)
    body = lines[2:2 + 20480] + [f"{el} {x:.5f} {y:.5f} {z:.5f}" for mol in water for el, (x, y, z) in zip("OHH", mol)]
    natoms = 20480 + 3 * nwater
    one = f"{natoms}\ncharge_arm fixture\n" + "\n".join(body) + "\n"
    with open(traj, "w") as f:
        f.write(one * 11)
    types = [(-1, 15, 512), (+1, 25, 512), (0, 3, nwater)]
    charges = []
    for ch, na, _ in types[:2]:
        q = rng.integers(-12, 13, size=na) / 16.0
        q[-1] += ch - q.sum()
        charges.append(q)
    charges.append(np.array([-0.8125, 0.40625, 0.40625]))
    jobs = [
        ("ca1", "charge_arm", [(Z, 1, 1), ((1, 1, 0), 2, 2), ((0, 1, 0), 3, 3)],
         dict(maxdist_optimize=True, sampling_interval=1, bin_count_a=30, bin_count_b=20)),
        ("ca2", "charge_arm", [((0, 0, -1), 2, 2), ((2, -1, 1), 1, 1), (Z, 3, 3)],
         dict(maxdist_optimize=True, normalise_CLM=True, subtract_uniform=True, sampling_interval=1, bin_count_a=12, bin_count_b=9)),
        ("ca3", "charge_arm", [(Z, 2, 2)], dict(maxdist="13.1", sampling_interval=5, bin_count_a=13, bin_count_b=1)),
    ]
    res = run_case("charge_arm", traj, "xyz", types, 11, ("0.0", "63.9223"), jobs, natoms, atom_charges=charges)
    shutil.rmtree(tmp)
    lo, hi = np.float32(0.0), np.float32(63.9223)
    save("charge_arm", res, types, [lo] * 3, [hi] * 3, jobs, res["coords"][:1], res["elements"],
         extra=dict(nsteps=np.int64(11), replicate=np.int64(11), atom_charge_0=charges[0], atom_charge_1=charges[1],
                    atom_charge_2=charges[2]))


CLUSTER_INPUTS = {
    # the two inputs Research_Example_4/general.inp runs (cluster_PAIRS.inp / cluster_GLOBAL.inp), 12 steps
    "cluster_PAIRS.inp": "PAIRS\n\nnsteps 12\nsampling_interval 1\nprint_spectators T\nprint_statistics T\nprint_step 1\n"
                         "pair 1 1 3 1 4.25 P-Li\npair 2 2 3 1 2.85 O-Li\npair 2 5 3 1 2.85 O-Li\nquit\n",
    "cluster_GLOBAL.inp": "GLOBAL\n\nnsteps 12\nsampling_interval 1\nprint_spectators T\nprint_statistics T\nprint_step 1\n"
                          "add_atom_index 1 1 # P atom\nadd_atom_index 3 1 # Li atom\nsingle_cutoff 4.25 #use this cutoff\nquit\n",
    # several cutoffs (scan + single, unsorted), custom weights, a whole molecule type, a sampling interval
    "cluster_SCAN.inp": "GLOBAL\nnsteps 11\nsampling_interval 2\nprint_statistics T\nautocorrelation F\ncustom_weight 0.25\n"
                        "add_molecule_type 1\ncustom_weight -1.5\nadd_atom_index 3 1\nsingle_cutoff 3.1\nscan_cutoff 1.5 2.75 3\nquit\n",
    # pairs that share lists in every way the reader distinguishes (A=old A, A=old B, B=old A, B=new A), custom weights
    "cluster_PAIRS2.inp": "pairs\nnsteps 12\nsampling_interval 3\nprint_statistics T\nautocorrelation F\ncustom_weight 2.0\n"
                          "pair 3 1 1 2 2.3\ncustom_weight -0.5\npair 1 2 2 2 3.4\npair 2 5 3 1 2.9\npair 2 2 2 2 3.3\npair 1 3 1 3 3.0\nquit\n",
}


def case_re4_cluster():
    """SURVEY 8f-4 consumer: the cluster analysis (Module_Cluster) of the reference binary on the first 12 frames of
    Research_Example_4 -- the `*cluster_statistics.dat` files it writes for the inputs above (the coordinates are those
    of re4.npz)."""
    src = os.path.join(REF, "Research_Example_4", "LiG1G1G1DFP-noH-1ns-363K.xyz")
    wd = tempfile.mkdtemp()
    with open(src) as f, open(os.path.join(wd, "traj.xyz"), "w") as g:
        for _ in range(12 * (1680 + 2)):
            g.write(f.readline())
    os.makedirs(os.path.join(wd, "output"))
    open(os.path.join(wd, "molecular.inp"), "w").write(" 12\n 3\n -1 5 70\n +0 6 210\n +1 1 70\nquit\n")
    gen = ' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./output/"\n set_threads 4\n cubic_box_edge 0.0 34.7853\n parallel_operation T\n sequential_read F\n trajectory_type xyz\n'
    for name, text in CLUSTER_INPUTS.items():
        open(os.path.join(wd, name), "w").write(text)
        gen += f' set_prefix "{name[8:-4]}_"\n cluster {name}\n'
    open(os.path.join(wd, "general.inp"), "w").write(gen + " quit\n")
    out, _ = orc.run_reference(wd)
    files = {}
    for name in CLUSTER_INPUTS:
        fn = f"{name[8:-4]}_cluster_statistics.dat"
        files[fn] = open(os.path.join(wd, "output", fn), "rb").read()
    ccf = [ln for ln in out.splitlines() if "CCF" in ln or (ln.strip() and ln.strip()[0].isdigit() and len(ln.split()) == 2)]
    np.savez_compressed(os.path.join(OUT, "re4_cluster.npz"),
                        input_names=np.array(list(CLUSTER_INPUTS)), input_texts=np.array(list(CLUSTER_INPUTS.values())),
                        file_names=np.array(list(files)), file_bytes=np.array([np.frombuffer(b, dtype=np.uint8) for b in files.values()], dtype=object),
                        ccf_lines=np.array(ccf), stdout=np.array(out))
    shutil.rmtree(wd)
    print("re4_cluster:", {k: len(v) for k, v in files.items()})


SPECIATION_INPUTS = {
    # Research_Example_4's own two inputs on 12 frames (nsteps / tmax shortened), autocorrelation and dumps off
    "spec_toLi.inp": " 1 acceptor\n 3 1 0.0\n 3 donors\n 1 1 4.25 to phosphorus\n 2 2 3.0\n 2 5 3.0\n new_donor_group 2\n assign_to_donor_group 2\n assign_to_donor_group 5\n"
                     " N_species 50\n N_neighbours 40\n nsteps 12\n tmax 11\n use_ccc T\n sampling_interval 1\n quit\n",
    "spec_fromLi.inp": " 3 acceptors\n 1 1 4.25 to phosphorus\n 2 2 3.0\n 2 5 3.0\n 1 donor\n 3 1 0.0\n new_acceptor_group 2\n assign_to_acceptor_group 2\n assign_to_acceptor_group 5\n"
                       " N_species 40\n N_neighbours 40\n nsteps 12\n tmax 11\n use_ccc T\n sampling_interval 1\n quit\n",
    # every second step, radii on both sides (cutoff = acceptor + donor contribution), fluorine and carbon atoms as well
    "spec_mixed.inp": " 2 acceptors\n 3 1 1.0\n 1 1 2.0\n 4 donors\n 2 2 1.9\n 2 5 2.0\n 1 4 1.25\n 2 1 2.5\n"
                      " N_species 60\n N_neighbours 40\n nsteps 12\n tmax 11\n use_ccc T\n sampling_interval 2\n quit\n",
    # limits: species overflow (error 161) and connection overflow (error 160), one grouped and one ungrouped acceptor
    # atom in the same molecule, generous radii so that acceptor molecules share donors
    "spec_limits.inp": " 3 acceptors\n 2 2 1.8\n 2 5 1.8\n 1 1 2.2\n 3 donors\n 3 1 1.7\n 1 2 1.2\n 1 3 1.2\n new_acceptor_group 2\n assign_to_acceptor_group 2\n"
                       " new_donor_group 1\n assign_to_donor_group 2\n assign_to_donor_group 3\n"
                       " N_species 7\n N_neighbours 3\n nsteps 11\n tmax 5\n use_ccc T\n sampling_interval 3\n quit\n",
    # no communal cluster correction (no table file), no groups, several atoms per acceptor and donor
    "spec_plain.inp": " 2 acceptors\n 2 2 1.5\n 2 5 1.5\n 3 donors\n 3 1 1.5\n 1 4 1.6\n 1 5 1.6\n"
                      " N_species 30\n N_neighbours 12\n nsteps 8\n tmax 5\n sampling_interval 1\n quit\n",
}


def _parse_speciation_statistics(text):
    """`*_speciation_statistics.dat` -> list of (percent string, connections, neighbour atoms, neighbour molecules)."""
    out, cur = [], None
    for ln in text.splitlines():
        t = ln.strip()
        if t.startswith("Species #("):
            cur = [t.split(" in ")[1].split("%")[0], None, None, None]
            out.append(cur)
        elif t.startswith("This species has a total of") and cur is not None:
            cur[1] = int(t.split()[6])
        elif t.startswith("with a total of") and cur is not None:
            w = t.split()
            cur[2], cur[3] = int(w[4]), int(w[8])
    return out


def case_re4_speciation():
    """SURVEY 8f-4, the speciation module's neighbour stage (Speciation:1405-1500): the reference binary's speciation
    analysis of the first 12 frames of Research_Example_4.  The species it reports (occurrence fraction <h> to 10 digits
    from `*_speciation_statistics_table.dat`, connections / neighbour atoms / neighbour molecules from
    `*_speciation_statistics.dat`) fix, for every acceptor molecule type, how many (acceptor molecule, step) cases have
    which (connections, neighbour atoms, neighbour molecules) triple -- a function of the neighbour test alone."""
    src = os.path.join(REF, "Research_Example_4", "LiG1G1G1DFP-noH-1ns-363K.xyz")
    wd = tempfile.mkdtemp()
    with open(src) as f, open(os.path.join(wd, "traj.xyz"), "w") as g:
        for _ in range(12 * (1680 + 2)):
            g.write(f.readline())
    os.makedirs(os.path.join(wd, "output"))
    open(os.path.join(wd, "molecular.inp"), "w").write(" 12\n 3\n -1 5 70\n +0 6 210\n +1 1 70\nquit\n")
    gen = ' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./output/"\n cubic_box_edge 0.0 34.7853\n parallel_operation F\n sequential_read F\n trajectory_type xyz\n'
    for name, text in SPECIATION_INPUTS.items():
        open(os.path.join(wd, name), "w").write(text)
        gen += f' set_prefix "{name[5:-4]}_"\n speciation "{name}"\n'
    open(os.path.join(wd, "general.inp"), "w").write(gen + " quit\n")
    out, _ = orc.run_reference(wd)
    names, texts, tables, none_lines = [], [], [], []
    for fn in sorted(os.listdir(os.path.join(wd, "output"))):
        if fn.endswith("_speciation_statistics.dat"):
            names.append(fn)
            texts.append(open(os.path.join(wd, "output", fn)).read())
            tp = os.path.join(wd, "output", fn[:-4] + "_table.dat")
            tables.append(open(tp).read() if os.path.exists(tp) else "")
    keep = [ln for ln in out.splitlines() if "without neighbouring donors" in ln or "different species were observed" in ln or "For this acceptor" in ln
            or "There was one acceptor" in ln or "Acceptor " in ln or "overflow" in ln.lower()]
    np.savez_compressed(os.path.join(OUT, "re4_speciation.npz"),
                        input_names=np.array(list(SPECIATION_INPUTS)), input_texts=np.array(list(SPECIATION_INPUTS.values())),
                        file_names=np.array(names), file_texts=np.array(texts), table_texts=np.array(tables),
                        stdout_lines=np.array(keep), stdout=np.array(out))
    shutil.rmtree(wd)
    print("re4_speciation:", names, [len(_parse_speciation_statistics(t)) for t in texts])
    print("\n".join(keep))


CASES = {"re4_speciation": case_re4_speciation, "lmp_masses": case_lmp_masses, "re4_cluster": case_re4_cluster, "dist_simple": case_dist_simple, "contact": case_contact, "charge_arm": case_charge_arm, "lmp": case_lmp, "dist": case_dist, "bmim": case_bmim, "re4": case_re4, "re4_atoms": case_re4_atoms, "synth": case_synth}

if __name__ == "__main__":
    assert orc.ref_available(), "run `make -C oracle` first (needs /root/reference)"
    for c in (sys.argv[1:] or list(CASES)):
        CASES[c]()
