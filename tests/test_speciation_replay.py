"""SURVEY 8f-4 consumer: the `speciation` keyword of the drop-in (Module_Speciation's statistics path) against the files
the reference binary wrote for three inputs on the first 12 frames of Research_Example_4
(tests/golden/re4_speciation.npz): `*_speciation_statistics.dat` and `*_speciation_statistics_table.dat`, byte for byte.

CPU suite: the host-side bookkeeping (species clipboard, atom groups, greedy recognition of known species, communal
cluster correction, writers) through the test seam pa_debug_run_general_input_with_connections, which takes the
connections of every step from a table -- filled here from the ORACLE's neighbour loops -- instead of the GPU primitive.
GPU suite: the real thing, `prealpha_b200_cli general.inp` with the connections from pa_neighbour_pairs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import fixtures
from prealpha_b200 import api, capi
from test_speciation_neighbours import _golden, _parse_input, _types_in_order


def _write_case(wd, fx, inputs, oracle):
    (wd / "output").mkdir(parents=True)
    oracle.write_xyz(str(wd / "traj.xyz"), fx.elements, fx.coords.astype(np.float64), fmt="%.9g")
    (wd / "molecular.inp").write_text(f" {fx.nsteps}\n {len(fx.types)}\n" + "".join(f" {c:+d} {a} {m}\n" for c, a, m in fx.types) + "quit\n")
    gen = ' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./output/"\n cubic_box_edge 0.0 34.7853\n parallel_operation F\n sequential_read F\n trajectory_type xyz\n'
    for name, text in inputs.items():
        (wd / name).write_text(text)
        gen += f' set_prefix "{name[5:-4]}_"\n speciation "{name}"\n'
    (wd / "general.inp").write_text(gen + " quit\n")


def _compare(wd, files):
    for fn, (stat, table) in files.items():
        got = (wd / "output" / fn).read_text()
        assert got == stat, (fn, _first_difference(got, stat))
        tp = wd / "output" / (fn[:-4] + "_table.dat")
        if table:
            got = tp.read_text()
            assert got == table, (fn, _first_difference(got, table))
        else:
            assert not tp.exists()


def _first_difference(a, b):
    la, lb = a.splitlines(), b.splitlines()
    for k, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return k, x, y
    return len(la), len(lb)


def test_speciation_bookkeeping_reproduces_the_reference_files(oracle, tmp_path):
    inputs, files = _golden()
    fx = fixtures.Fixture("re4")
    traj = fx.trajectory()
    wd = tmp_path / "run"
    _write_case(wd, fx, inputs, oracle)
    rows = []
    for ordinal, (name, text) in enumerate(inputs.items()):
        acc, don, nsteps, dt = _parse_input(text)
        don_types = _types_in_order(don)
        b, b_cls, b_meta = [], [], []
        for nd, t in enumerate(don_types, start=1):
            entries = [(e, at) for e, (t2, at, _) in enumerate(don) if t2 == t]
            for m in range(1, fx.types[t - 1][2] + 1):
                for k, (e, at) in enumerate(entries, start=1):
                    b.append((t, m, at)); b_cls.append(e); b_meta.append((nd, m, k))
        b = np.array(b, dtype=np.int32); b_cls = np.array(b_cls, dtype=np.int32)
        cut = np.array([[(ra + rd) ** 2 for _, _, rd in don] for _, _, ra in acc], dtype=np.float32)
        for n_acc, ty in enumerate(_types_in_order(acc), start=1):
            entries = [(e, at) for e, (t2, at, _) in enumerate(acc) if t2 == ty]
            a, a_cls, a_meta = [], [], []
            for m in range(1, fx.types[ty - 1][2] + 1):
                for k, (e, at) in enumerate(entries, start=1):
                    a.append((ty, m, at)); a_cls.append(e); a_meta.append((m, k))
            a = np.array(a, dtype=np.int32); a_cls = np.array(a_cls, dtype=np.int32)
            for step in range(1, nsteps + 1, dt):
                rq, keep = capi.make_neigh_request(step, a=a, b=b, cutoff_sq=cut, a_class=a_cls, b_class=b_cls, skip_same_molecule=True)
                for ia, ib in oracle.neighbour_pairs(traj, rq):
                    rows.append((ordinal, step, n_acc, a_meta[ia][0], b_meta[ib][0], b_meta[ib][1], b_meta[ib][2], a_meta[ia][1]))
    table = np.ascontiguousarray(np.array(rows, dtype=np.int32))
    np.random.default_rng(0).shuffle(table)                      # the seam orders the rows itself
    table = np.ascontiguousarray(table)
    lib = capi.load()
    fn = lib.pa_debug_run_general_input_with_connections
    fn.restype = C.c_int
    fn.argtypes = [C.c_char_p, C.POINTER(capi.pa_exec), C.POINTER(C.c_int32), C.c_int64]
    cwd = os.getcwd()
    os.chdir(wd)
    try:
        rc = fn(b"general.inp", C.byref(api.make_exec()), table.ctypes.data_as(C.POINTER(C.c_int32)), len(table))
    finally:
        os.chdir(cwd)
    # as the reference: warning 55 (xyz trajectory), 28 (tmax beyond the sampled steps) in spec_mixed and spec_limits, 160
    # (connection overflow) in spec_limits; the species-overflow NOTICEs 161 are reported but not counted
    assert rc == 4
    _compare(wd, files)


@pytest.mark.gpu
@pytest.mark.parametrize("brute", [False, True], ids=["cell_list", "brute_force"])
def test_speciation_keyword_on_the_gpu_reproduces_the_reference_files(brute, oracle, tmp_path):
    assert api.device_count() >= 1, "GPU tests need a B200; the product has no CPU fallback"
    inputs, files = _golden()
    fx = fixtures.Fixture("re4")
    wd = tmp_path / "run"
    _write_case(wd, fx, inputs, oracle)
    cli = os.path.join(os.path.dirname(capi.LIB_PATH), "prealpha_b200_cli")
    env = dict(os.environ, PA_NEIGH_MIN_CELLS="100000000" if brute else "1")
    p = subprocess.run([cli, "general.inp"], cwd=wd, capture_output=True, timeout=600, env=env)
    out = p.stdout.decode()
    # exit status 0 like the binary's; the report ends with the reference's count (55, 28, 28, 160)
    assert p.returncode == 0 and "Encountered 4 errors/warnings globally." in out, out[-3000:] + p.stderr.decode()[-2000:]
    assert "B200 execution (speciation)" in out
    _compare(wd, files)
