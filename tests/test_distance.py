"""Row X (Module_Distances): oracle against the reference binary's report (CPU), CUDA against the oracle
and the CLI against the binary's stdout (GPU)."""
import os
import re
import subprocess

import numpy as np
import pytest

import fixtures
from prealpha_b200 import api, capi, inputs


def _dist_fixture():
    z = np.load(os.path.join(fixtures.GOLDEN, "dist.npz"))
    return ([str(s) for s in z["input_names"]], [str(s) for s in z["input_texts"]], z["stdout"].tobytes().decode())


def _bmim_two_steps():
    fx = fixtures.Fixture("bmim")
    fx.coords = fx.coords[:2]
    fx.nsteps = 2
    fx.atom_charges = None
    return fx, fx.trajectory()


def _type_info(fx):
    els, nmol, off = [], [], 0
    for ch, na, nm in fx.types:
        els.append(fx.elements[off:off + na]); nmol.append(nm); off += na * nm
    return els, nmol


def _ref_report_blocks(stdout):
    """Lines between 'Done with ...molecular distance calculation. Results:' and ' ( reference atom(s)'."""
    blocks, cur = [], None
    for line in stdout.splitlines():
        if "molecular distance calculation. Results:" in line:
            cur = []
        elif line.startswith(" ( reference atom(s)") and cur is not None:
            blocks.append(cur); cur = None
        elif cur is not None:
            cur.append(line.rstrip())
    return blocks


def test_distance_oracle_matches_reference_binary_report(oracle):
    names, texts, stdout = _dist_fixture()
    fx, traj = _bmim_two_steps()
    els, nmol = _type_info(fx)
    blocks = _ref_report_blocks(stdout)
    assert len(blocks) == len(texts) == 3
    sticky = {}
    for text, ref_lines in zip(texts, blocks):
        spec = inputs.parse_distance_input(text, els, nmol, box_limit=float(fx.box_hi[0] - fx.box_lo[0]), sticky=sticky)
        job, keep = inputs.make_dist_job(spec)
        res = oracle.distance_subjobs(traj, job)
        assert inputs.distance_report(spec, res) == ref_lines


@pytest.mark.gpu
def test_distance_cuda_matches_oracle(oracle):
    """Bit-exact: the kernel returns the squared distances below the cutoff in the reference's loop order and the host
    replays the reference's sequential REAL(8) accumulation on them (same chain of additions, same libm expf)."""
    names, texts, stdout = _dist_fixture()
    fx, traj = _bmim_two_steps()
    els, nmol = _type_info(fx)
    sticky = {}
    blocks = _ref_report_blocks(stdout)
    for text, ref_lines in zip(texts, blocks):
        spec = inputs.parse_distance_input(text, els, nmol, box_limit=float(fx.box_hi[0] - fx.box_lo[0]), sticky=sticky)
        job, keep = inputs.make_dist_job(spec)
        got = api.distance_subjobs(traj, job)
        exp = oracle.distance_subjobs(traj, job)
        for g, e in zip(got, exp):
            assert g["missed_neighbours"] == e["missed_neighbours"]
            assert g["ffc_weight"] == e["ffc_weight"]
            for key in ("closest_average", "closest_stdev", "exponential_average", "exponential_stdev", "exponential_weight",
                        "ffc_average", "ffcexp_average"):
                assert g[key] == e[key] or (g[key] != g[key] and e[key] != e[key]), key      # identical doubles (NaN == NaN)
        assert inputs.distance_report(spec, got) == ref_lines


@pytest.mark.gpu
def test_distance_cli_matches_reference_stdout(oracle, tmp_path):
    names, texts, stdout = _dist_fixture()
    fx, _ = _bmim_two_steps()
    wd = tmp_path / "run"
    (wd / "out").mkdir(parents=True)
    oracle.write_xyz(str(wd / "traj.xyz"), fx.elements, fx.coords.astype(np.float64), fmt="%.4f")
    (wd / "molecular.inp").write_text(" 2\n 2\n -1 15 512\n +1 25 512\nquit\n")
    gen = ' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./out/"\n cubic_box_edge 0.0 63.9223\n sequential_read F\n'
    for n, t in zip(names, texts):
        (wd / n).write_text(t)
        gen += f' distance "{n}"\n'
    (wd / "general.inp").write_text(gen + " quit\n")
    cli = os.path.join(os.path.dirname(capi.LIB_PATH), "prealpha_b200_cli")
    p = subprocess.run([cli, "general.inp"], cwd=wd, capture_output=True, timeout=600)
    out = p.stdout.decode()
    assert _ref_report_blocks(out) == _ref_report_blocks(stdout), out[-3000:]


# ---- contact_distance (Debug:586-623) -------------------------------------------------------------------------------

_CONTACT_LINE = re.compile(r"^\s*(Calculating intra- and intermolecular contact distances.*|Molecule type index \d+ out of \d+\."
                           r"|(?:Largest|Smallest) int(?:ra|er)molecular distance:\s+\S+)\s*$")


def _contact_lines(stdout):
    return [m.group(1) for m in (_CONTACT_LINE.match(l) for l in stdout.splitlines()) if m]


def _contact_fixture():
    z = np.load(os.path.join(fixtures.GOLDEN, "contact.npz"))
    return [str(k) for k in z["keywords"]], z["stdout"].tobytes().decode()


def _f0_3(v):
    t = "%.3f" % float(v)
    return t[1:] if t.startswith("0.") else t


def _contact_calls(keywords, nsteps, ntypes):
    """(timestep, type) per print_distances call, in order, with check_timestep applied (Debug:46-56)."""
    calls = []
    for kw in keywords:
        tk = kw.split()
        step, ty = (1, -1) if tk[0] == "contact_distance_simple" else (int(tk[1]), int(tk[2]))
        step = min(max(step, 1), nsteps)
        calls += [(step, t) for t in (range(1, ntypes + 1) if ty < 1 else [ty])]
    return calls


def _contact_numbers(stdout):
    return [l.split(":")[1].strip() for l in _contact_lines(stdout) if "distance:" in l]


def test_contact_distance_oracle_matches_reference_binary(oracle):
    keywords, stdout = _contact_fixture()
    fx = fixtures.Fixture("lmp")
    traj = fx.trajectory()
    got = []
    for step, ty in _contact_calls(keywords, fx.nsteps, len(fx.types)):
        got += [_f0_3(v) for v in oracle.contact_distance(traj, step, ty)]
    assert got == _contact_numbers(stdout) and len(got) == 3 * 11


@pytest.mark.gpu
def test_contact_distance_cuda_matches_oracle_bitwise(oracle):
    keywords, stdout = _contact_fixture()
    fx = fixtures.Fixture("lmp")
    traj = fx.trajectory()
    got = []
    for step, ty in sorted(set(_contact_calls(keywords, fx.nsteps, len(fx.types)))):
        res, pairs = api.contact_distance(traj, step, ty)
        exp = oracle.contact_distance(traj, step, ty)
        assert [np.float32(v).tobytes() for v in res] == [np.float32(v).tobytes() for v in exp], (step, ty)
        assert pairs > 0
    # a larger system with one-atom types, and the bmim frame (many atoms per molecule)
    rng = np.random.default_rng(77)
    types = []
    for ch, na, nm in [(1, 1, 3000), (-1, 1, 2500), (0, 3, 700)]:
        c = rng.random((1, nm, 1, 3)) * 40.0 - 5.0
        types.append(dict(charge=ch, natoms=na, nmol=nm, elements=["C"] * na,
                          xyz=np.round(c + rng.normal(0, 0.5, size=(1, nm, na, 3)), 4).astype(np.float32)))
    big = capi.Trajectory(types, [0.0] * 3, [30.0, 31.5, 29.25])
    for ty in (1, 2, 3):
        res, _ = api.contact_distance(big, 1, ty)
        exp = oracle.contact_distance(big, 1, ty)
        assert [np.float32(v).tobytes() for v in res] == [np.float32(v).tobytes() for v in exp], ty
    with pytest.raises(RuntimeError):
        api.contact_distance(big, 2, 1)          # only one step loaded
    with pytest.raises(RuntimeError):
        api.contact_distance(big, 1, 4)


@pytest.mark.gpu
def test_contact_distance_cli_matches_reference_stdout(oracle, tmp_path):
    keywords, stdout = _contact_fixture()
    fx = fixtures.Fixture("lmp")
    wd = tmp_path / "run"
    (wd / "out").mkdir(parents=True)
    oracle.write_lmp(str(wd / "traj.lmp"), fx.elements, fx.coords.astype(np.float64), fx.box_lo, fx.box_hi, fmt="%.9g")
    oracle.write_case(str(wd), "traj.lmp", fx.types, fx.nsteps, None, {}, extra_general=["trajectory_type lmp"] + keywords)
    cli = os.path.join(os.path.dirname(capi.LIB_PATH), "prealpha_b200_cli")
    p = subprocess.run([cli, "general.inp"], cwd=wd, capture_output=True, timeout=600)
    out = p.stdout.decode()
    assert p.returncode >= 0, out[-3000:]          # not killed by a signal; warnings make the exit status non-zero
    assert _contact_lines(out) == _contact_lines(stdout), out[-3000:]
    assert "Encountered 2 errors/warnings globally." in out


@pytest.mark.gpu
def test_distance_simple_cli_matches_reference_stdout(oracle, tmp_path):
    """`distance_simple` (Main:3394-3409): the generated prealpha_simple.inp files (intra, then inter with the
    reference's F-only `both_available` test) and both report blocks equal the binary's."""
    z = np.load(os.path.join(fixtures.GOLDEN, "dist_simple.npz"))
    stdout, simple = z["stdout"].tobytes().decode(), str(z["simple_inp"])
    fx, _ = _bmim_two_steps()
    wd = tmp_path / "run"
    (wd / "out").mkdir(parents=True)
    oracle.write_xyz(str(wd / "traj.xyz"), fx.elements, fx.coords.astype(np.float64), fmt="%.4f")
    oracle.write_case(str(wd), "traj.xyz", fx.types, 2, ("0.0", "63.9223"), {}, extra_general=["distance_simple"])
    cli = os.path.join(os.path.dirname(capi.LIB_PATH), "prealpha_b200_cli")
    p = subprocess.run([cli, "general.inp"], cwd=wd, capture_output=True, timeout=900)
    out = p.stdout.decode()
    blocks = _ref_report_blocks(out)
    assert len(blocks) == 2 and blocks == _ref_report_blocks(stdout), out[-3000:]
    assert (wd / "prealpha_simple.inp").read_text() == simple          # the last one written: intermolecular
