"""The CPU restatement (oracle/pa_oracle.c) against what the reference's own binary produced
(tests/golden/*.npz): `within bin` / `out of bounds` integers and every output file byte for byte."""
import os

import pytest

import fixtures


@pytest.mark.parametrize("name", fixtures.available())
def test_oracle_matches_reference_binary(name, oracle, tmp_path):
    fx = fixtures.Fixture(name)
    traj = fx.trajectory()
    for jn, refnum, rq, (within, oob) in fx.jobs():
        job = oracle.job_setup(traj, rq)
        hist, w, o = oracle.histogram(traj, job)
        assert (w, o) == (within, oob), (name, jn, refnum)
        assert int(hist.sum()) == w or job.weigh_charge
        g, rho, rc = oracle.transfer(traj, job, hist, w, o, verbose=1)
        assert rc == 0
        outdir = tmp_path / f"{jn}_{refnum}"
        outdir.mkdir()
        oracle.write(traj, job, refnum, str(outdir) + os.sep, jn + ".inp", g, rho)
        exp = fx.expected_files(jn)
        produced = sorted(os.listdir(outdir))
        assert produced, (jn, refnum)
        for fn in produced:
            assert fn in exp, (fn, sorted(exp))
            got = (outdir / fn).read_bytes()
            assert got == exp[fn], f"{name}/{jn}/{fn} differs from the reference binary's output"


def test_every_reference_file_is_covered(oracle, tmp_path):
    """No output file of the binary is left unchecked (file-name logic of Distribution:999-1059)."""
    for name in fixtures.available():
        fx = fixtures.Fixture(name)
        traj = fx.trajectory()
        made = set()
        for jn, refnum, rq, _ in fx.jobs():
            job = oracle.job_setup(traj, rq)
            hist, w, o = oracle.histogram(traj, job)
            g, rho, _ = oracle.transfer(traj, job, hist, w, o)
            d = tmp_path / name / f"{jn}_{refnum}"
            d.mkdir(parents=True)
            oracle.write(traj, job, refnum, str(d) + os.sep, jn + ".inp", g, rho)
            made |= {jn + "_" + f for f in os.listdir(d)}
        assert made == set(fx.files), (sorted(set(fx.files) - made), sorted(made - set(fx.files)))
