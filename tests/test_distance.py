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
    """fp64 sums are reduced in a fixed tree order on the GPU and in loop order by the reference:
    tolerance 1e-9 relative (the reference prints 3-4 significant digits); closest distances are exact."""
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
                assert g[key] == pytest.approx(e[key], rel=1e-9, abs=1e-300), key
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
