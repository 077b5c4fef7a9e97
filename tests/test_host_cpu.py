"""CPU-side checks of the product library (no compute calls): the C ABI loads and exports every
declared symbol, and the host logic around the hot loop (job set-up, class tables, normalisation,
writers) reproduces the reference binary's outputs when fed the oracle's histogram."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fixtures
from prealpha_b200 import api, capi, inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    header = open(os.path.join(ROOT, "include", "prealpha_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(capi.SIGNATURES), (declared ^ set(capi.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pa_abi_version() == 2


def test_struct_layout_matches_header_sizes():
    # sizes the C compiler produced (pa_debug_* is built from the same header)
    assert C.sizeof(capi.pa_type) == 56
    assert C.sizeof(capi.pa_traj) == 16 + 48 + 8 + 8
    assert C.sizeof(capi.pa_job) == 4 * 6 + 24 + 72 + 16 + 8 * 4
    assert C.sizeof(capi.pa_exec) == 32


def test_ctypes_structs_match_the_header_as_compiled_by_gcc(tmp_path):
    """Every struct of include/prealpha_b200.h: sizeof and the offset of its last field, printed by a C program compiled
    against the header, equal what the ctypes mirror in prealpha_b200/capi.py lays out."""
    import subprocess
    structs = {"pa_type": "xyz", "pa_traj": "wrap_trajectory", "pa_job_request": "normalise_clm", "pa_job": "normalise_clm",
               "pa_exec": None, "pa_stats": None, "pa_dist_subjob": "weight", "pa_dist_job": "exponent", "pa_dist_result": "missed_neighbours",
               "pa_contact_result": "pairs_evaluated", "pa_neigh_request": "skip_same_molecule", "pa_neigh_pair": "ib"}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "prealpha_b200.h"', "int main(void) {"]
    for name, last in structs.items():
        off = f"(long)offsetof({name}, {last})" if last else "-1L"
        lines.append(f'  printf("{name} %zu %ld\\n", sizeof({name}), {off});')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call(["gcc", "-std=c11", "-I", inc, str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    seen = 0
    for row in out:
        if not row.strip():
            continue
        name, size, off = row.split()
        ct = getattr(capi, name)
        assert C.sizeof(ct) == int(size), name
        if int(off) >= 0:
            assert getattr(ct, structs[name]).offset == int(off), name
        seen += 1
    assert seen == len(structs)


@pytest.mark.parametrize("name", fixtures.available())
def test_job_setup_matches_oracle_bitwise(name, oracle):
    fx = fixtures.Fixture(name)
    traj = fx.trajectory()
    for jn, refnum, rq, _ in fx.jobs():
        a = api.job_setup(traj, rq)
        b = oracle.job_setup(traj, rq)
        assert bytes(a) == bytes(b), (name, jn, refnum)


@pytest.mark.parametrize("name", fixtures.available())
def test_transfer_and_writers_reproduce_reference_files(name, oracle, tmp_path):
    """Product host code (rows N, W) on the oracle's integer histogram -> the binary's files."""
    fx = fixtures.Fixture(name)
    traj = fx.trajectory()
    for jn, refnum, rq, (within, oob) in fx.jobs():
        job = api.job_setup(traj, rq)
        hist, w, o = oracle.histogram(traj, job)
        assert (w, o) == (within, oob)
        g, rho, rc = api.transfer(traj, job, hist, w, o, verbose=1)
        assert rc == 0
        g0, rho0, _ = oracle.transfer(traj, job, hist, w, o, verbose=1)
        assert g.tobytes() == g0.tobytes() and np.float32(rho).tobytes() == np.float32(rho0).tobytes()
        d = tmp_path / f"{jn}_{refnum}"
        d.mkdir()
        api.write(traj, job, refnum, str(d) + os.sep, jn + ".inp", g, rho)
        exp = fx.expected_files(jn)
        for fn in os.listdir(d):
            assert (d / fn).read_bytes() == exp[fn], (name, jn, fn)


def test_format_real4_matches_oracle(oracle):
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.standard_normal(2000).astype(np.float32) * np.float32(10.0) ** rng.integers(-12, 12, 2000).astype(np.float32),
        np.array([0.0, 0.1, 0.099999994, 0.09999999, 1e9, 999999999.0, 9.99999999e8, 1.0, -1.0, 123456.789, 1e-38, 3e38,
                  0.99999994, 9.9999999], dtype=np.float32)])
    for v in vals:
        assert api.format_real4(v) == oracle.format_real4(v), float(v)
    assert api.format_real4(0.159805745) == "  0.159805745    "
    assert api.format_real4(5.32685816e-02) == "   5.32685816E-02"


def _tables(job, mds):
    lib = capi.load()
    lib.pa_debug_thr_a.restype = C.c_int
    lib.pa_debug_thr_a.argtypes = [C.POINTER(capi.pa_job), C.c_double, C.POINTER(C.c_uint64), C.c_int32]
    lib.pa_debug_thr_b.restype = C.c_int
    lib.pa_debug_thr_b.argtypes = [C.POINTER(capi.pa_job), C.POINTER(C.c_double), C.c_int32]
    lib.pa_debug_a_class.restype = C.c_int
    lib.pa_debug_a_class.argtypes = [C.POINTER(capi.pa_job), C.c_double, C.c_double]
    lib.pa_debug_b_class.restype = C.c_int
    lib.pa_debug_b_class.argtypes = [C.POINTER(capi.pa_job), C.c_double]
    n = lib.pa_debug_thr_a(C.byref(job), mds, None, 0)
    ta = np.zeros(n, dtype=np.uint64)
    lib.pa_debug_thr_a(C.byref(job), mds, ta.ctypes.data_as(C.POINTER(C.c_uint64)), n)
    tb = None
    if job.mode == capi.PA_MODE_PDF:
        n = lib.pa_debug_thr_b(C.byref(job), None, 0)
        tb = np.zeros(n, dtype=np.float64)
        lib.pa_debug_thr_b(C.byref(job), tb.ctypes.data_as(C.POINTER(C.c_double)), n)
    return lib, ta, tb


@pytest.mark.parametrize("maxdist,bins,mds", [(31.96115, 300, 22500.0), (135.7209, 1000, 22500.0), (7.5, 64, 22500.0),
                                              (200.0, 500, 22500.0), (10.0, 1, 1e9), (0.5, 1000, 22500.0)])
def test_a_class_table_equals_reference_arithmetic(maxdist, bins, mds):
    """#thresholds <= bits(m)  ==  INT(SQRT(REAL4(m))/step_a) etc. for random and boundary m."""
    fx = fixtures.Fixture("synth")
    traj = fx.trajectory()
    rq = capi.make_request(mode="pdf", bin_a=bins, bin_b=1, maxdist=maxdist, sampling_interval=1)
    job = api.job_setup(traj, rq)
    lib, ta, _ = _tables(job, mds)
    assert len(ta) == bins + 3 and ta[0] == 0
    fin = ta[ta != np.uint64(0xFFFFFFFFFFFFFFFF)]
    assert np.all(np.diff(fin.astype(np.float64)) >= 0)
    rng = np.random.default_rng(5)
    ms = np.concatenate([rng.random(4000) * maxdist ** 2 * 1.3, rng.random(500) * 1e-6, [0.0, 1e300, np.inf, mds, np.nextafter(mds, 0)]])
    near = []
    for t in fin[1:]:
        v = np.array([t], dtype=np.uint64).view(np.float64)[0]
        near += [v, np.nextafter(v, 0.0), np.nextafter(v, np.inf)]
    ms = np.concatenate([ms, np.array(near)])
    bits = ms.view(np.uint64) if ms.dtype == np.float64 else None
    for m, b in zip(ms, np.asarray(ms, dtype=np.float64).view(np.uint64)):
        cls_table = int(np.searchsorted(ta[1:], b, side="right"))        # number of k>=1 with thr[k] <= bits
        assert cls_table == lib.pa_debug_a_class(C.byref(job), mds, float(m)), (m, cls_table)


@pytest.mark.parametrize("bins,vec", [(1, (0, 0, 1)), (36, (0, 0, 1)), (18, (1, 1, 0)), (1000, (2, -1, 1))])
def test_b_class_table_equals_acos_arithmetic(bins, vec):
    fx = fixtures.Fixture("synth")
    traj = fx.trajectory()
    job = api.job_setup(traj, capi.make_request(mode="pdf", vec=vec, bin_a=10, bin_b=bins, maxdist=5.0, sampling_interval=1))
    lib, _, tb = _tables(job, 22500.0)
    assert len(tb) == bins + 1
    assert np.all(np.diff(tb[1:]) <= 0)
    rng = np.random.default_rng(6)
    dots = np.concatenate([rng.random(3000) * 2 - 1, [1.0, -1.0, 0.0, -1 + 2.2e-16, -1 + 4.4e-16, -1 + 6.6e-16, 1 - 1.1e-16]])
    near = []
    for t in tb[1:]:
        if -1 <= t <= 1:
            near += [t, np.nextafter(t, -2.0), np.nextafter(t, 2.0)]
    dots = np.concatenate([dots, np.clip(np.array(near), -1, 1)])
    for d in dots:
        cls_table = int(np.sum(d <= tb[1:]))
        ref = lib.pa_debug_b_class(C.byref(job), float(d))
        assert cls_table == min(ref, bins), (d, cls_table, ref)


def test_compute_entry_points_fail_loudly_without_gpu():
    """No CPU fallback: without a device the hot path must return PA_ERR_NO_DEVICE, not a result."""
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    fx = fixtures.Fixture("synth")
    traj = fx.trajectory()
    job = api.job_setup(traj, fx.jobs()[0][2])
    hist = np.zeros(job.bin_a * job.bin_b, dtype=np.int64)
    w, o = C.c_int64(0), C.c_int64(0)
    rc = capi.load().pa_distribution_histogram(traj.ptr, C.byref(job), hist.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(w), C.byref(o))
    assert rc == capi.PA_ERR_NO_DEVICE
    assert not hist.any()
    assert b"no CUDA device" in capi.load().pa_last_error() or b"CPU fallback" in capi.load().pa_last_error()


def test_rdf_fast_eligibility():
    fx = fixtures.Fixture("synth")
    traj = fx.trajectory()
    lib = capi.load()
    lib.pa_debug_is_rdf_fast.restype = C.c_int
    lib.pa_debug_is_rdf_fast.argtypes = [C.POINTER(capi.pa_job)]

    def fast(**kw):
        return lib.pa_debug_is_rdf_fast(C.byref(api.job_setup(traj, capi.make_request(sampling_interval=1, **kw))))
    assert fast(mode="pdf", bin_b=1, maxdist=10.0) == 1
    assert fast(mode="pdf", vec=(0, 0, -3), bin_b=1, maxdist=10.0) == 1
    assert fast(mode="pdf", bin_b=2, maxdist=10.0) == 0
    assert fast(mode="pdf", vec=(1, 0, 0), bin_b=1, maxdist=10.0) == 0
    assert fast(mode="pdf", bin_b=1, maxdist=10.0, use_coc=True) == 0
    assert fast(mode="cdf", bin_b=1, maxdist=10.0) == 0


def test_distribution_input_parser():
    text = "podf 2\n+0 +0 +1 2 1\n 1, 1, 0, 1 -1 ### comment\nmaxdist 12.5\nbin_count 50\nbin_count_b 2000\nweigh_charge T\nunknown line\nquit\nbin_count_a 7\n"
    reqs = inputs.parse_distribution_input(text, ntypes=2)
    assert len(reqs) == 2
    assert reqs[0].mode == capi.PA_MODE_PDF and reqs[1].obs_type == -1 and list(reqs[1].ref_xyz) == [1, 1, 0]
    assert reqs[0].bin_a == 50 and reqs[0].bin_b == 1000 and reqs[0].weigh_charge == 1 and reqs[0].sampling_interval == 10
    assert reqs[0].maxdist == 12.5 and reqs[0].maxdist_optimize == 0
    with pytest.raises(ValueError):
        inputs.parse_distribution_input("pdf 1\n0 0 1 3 1\n", ntypes=2)


def test_fast_coordinate_parser_equals_strtof():
    """The trajectory readers convert decimal text to REAL(4) with a fast path (exact integer mantissa x exact power of ten
    in double, then one rounding to float32) that must agree with libc's correctly rounded strtof bit for bit, including
    how many characters it consumes: random decimals of every length, exponents, signs, zeros, strings on and next to
    float32 rounding midpoints (where the fast path must defer to strtof), overflow / underflow, and non-numeric tokens."""
    import ctypes.util
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    lib = capi.load()
    fn = lib.pa_debug_parse_coordinate
    fn.restype, fn.argtypes = C.c_float, [C.c_char_p, C.POINTER(C.c_int32)]
    libc = C.CDLL(ctypes.util.find_library("c"))
    libc.strtof.restype, libc.strtof.argtypes = C.c_float, [C.c_char_p, C.POINTER(C.c_char_p)]
    rng = np.random.default_rng(99)
    cases = ["0", "-0", "0.0", "-0.0", "+1", "1.", ".5", "-.5", ".", "-", "+", "", "1e", "1e+", "1E5", "1e-5", "1.5D0", "12,3", "7abc",
             "0x1p3", "0X10", "inf", "-inf", "nan", "NaN", "infinity", "1e38", "3.4028235e38", "3.4028236e38", "3.5e38", "1e39", "1e-37", "1e-38",
             "1.17549435e-38", "1e-40", "1e-45", "7e-46", "1e-50", "123456789012345", "1234567890123456", "0.000000000000000000001",
             "00000123.4500000", "1e22", "1e23", "1e-22", "1e-23", "9007199254740993", "16777217", "16777216.5", "33554433", "0.1", "0.3",
             "271.4418", "63.9223", "-12.34567", "1.0000000596046448", "1.00000005960464477539062", "1.00000005960464477539063", "1.00000005960464477539061"]
    for nd in range(1, 19):                                              # random digit strings with a point somewhere
        for _ in range(1500):
            digits = "".join(rng.choice(list("0123456789"), size=nd))
            k = int(rng.integers(0, nd + 1))
            s = digits[:k] + "." + digits[k:]
            if rng.random() < 0.3:
                s += f"e{int(rng.integers(-30, 31))}"
            cases.append(("-" if rng.random() < 0.5 else "") + s)
    for _ in range(4000):                                                # exact float32 midpoints and their decimal neighbours
        f = np.float32(10.0 ** rng.uniform(-6, 6) * rng.uniform(1, 10))
        g = np.nextafter(f, np.float32(np.inf), dtype=np.float32)
        mid = (Decimal(float(f)) + Decimal(float(g))) / 2
        txt = format(mid, "f")
        cases.append(txt)
        for digits in (9, 12, 15, 17):
            q = Decimal(1).scaleb(-digits)
            cases += [format(mid.quantize(q), "f"), format(mid.quantize(q) + q, "f"), format(mid.quantize(q) - q, "f")]
    for s in cases:
        b = s.encode()
        n = C.c_int32(-1)
        got = np.float32(fn(b, C.byref(n)))
        endp = C.c_char_p()
        buf = C.create_string_buffer(b)
        exp = np.float32(libc.strtof(buf, C.byref(endp)))
        consumed = C.cast(endp, C.c_void_p).value - C.addressof(buf)
        assert got.tobytes() == exp.tobytes() or (np.isnan(got) and np.isnan(exp)), (s, got, exp)
        assert n.value == consumed, (s, n.value, consumed)
    assert len(cases) > 40000


def _run_general_inp(workdir, env=None):
    """pa_run_general_input in a subprocess (its report goes to the C stdout); returns (error count, stdout)."""
    import subprocess, sys
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from prealpha_b200 import capi, api; "
            "rc = capi.load().pa_run_general_input(b'general.inp', C.byref(api.make_exec())); print('RC', rc)" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], cwd=workdir, capture_output=True, timeout=300,
                       env=dict(os.environ, **(env or {})))
    out = p.stdout.decode()
    assert p.returncode == 0, out + p.stderr.decode()
    return int(out.rsplit("RC", 1)[1]), out


def _write_small_case(wd, nsteps_declared=12, nsteps_written=12, natoms_line=None, body=""):
    rng = np.random.default_rng(3)
    n = 24
    with open(wd / "traj.xyz", "w") as f:
        for s in range(nsteps_written):
            f.write(f"{natoms_line if natoms_line is not None else n}\nstep {s}\n")
            for a in range(n):
                x = rng.random(3) * 12.0
                f.write(f"{'Na' if a < 12 else 'Cl'} {x[0]:.4f} {x[1]:.4f} {x[2]:.4f}\n")
    (wd / "molecular.inp").write_text(f" {nsteps_declared}\n 2\n +1 1 12\n -1 1 12\nquit\n")
    (wd / "general.inp").write_text(' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./"\n cubic_box_edge 0.0 12.0\n trajectory_type xyz\n'
                                    ' sequential_read F\n' + body + " quit\n")


def test_general_inp_host_logic_without_gpu(tmp_path):
    """The drop-in front end up to the point where a kernel would be needed: header, switches, molecular.inp, threaded
    trajectory reader, keyword dispatch and the reference's warning counter -- all host code, so it runs here."""
    wd = tmp_path / "ok"; wd.mkdir()
    _write_small_case(wd, body=" set_prefix test_\n gyradius -1 1 1\n this_keyword_does_not_exist 1\n #comment\n")
    rc, out = _run_general_inp(wd)
    assert "reading the trajectory...done (12 steps of 24 atoms" in out
    assert "belongs to a prealpha module outside the accelerated path" in out          # gyradius
    # (the count includes the reference's warning 55: an xyz trajectory does not say whether it holds positions or velocities)
    assert "ERROR 20" in out and "ERROR 55" in out and rc == 2 and "Encountered 2 errors/warnings globally." in out
    # fewer steps in the file than molecular.inp declares: error 84, the shorter trajectory is used (Molecular:5003-5009)
    wd = tmp_path / "short"; wd.mkdir()
    _write_small_case(wd, nsteps_declared=15, nsteps_written=12)
    rc, out = _run_general_inp(wd)
    assert "ERROR 84" in out and "using 12 steps" in out and rc == 2
    # atom count of the trajectory header differs from molecular.inp: error 10
    wd = tmp_path / "natoms"; wd.mkdir()
    _write_small_case(wd, natoms_line=25)
    rc, out = _run_general_inp(wd)
    assert "ERROR 10" in out
    # a compute keyword without a GPU must fail loudly (and only then), never fall back
    if api.device_count() == 0:
        wd = tmp_path / "compute"; wd.mkdir()
        (wd / "rdf.inp").write_text("pdf 1\n+0 +0 +1 1 2\nmaxdist_optimize\nsampling_interval 1\nbin_count_a 10\nbin_count_b 1\nquit\n")
        _write_small_case(wd, body=' distribution "rdf.inp"\n')
        rc, out = _run_general_inp(wd)
        assert "GPU histogram failed" in out and "no CPU fallback" in out and rc == capi.PA_ERR_CUDA      # a failed GPU call is not a warning count
        assert not any(f.startswith("reference_") for f in os.listdir(wd))


def test_binary_frame_cache_is_written_validated_and_reused(tmp_path):
    """PREALPHA_B200_CACHE: the first read of a text trajectory leaves the parsed float32 frames next to it, the second
    run reads them back (no text parsing), a changed source file or molecule table invalidates the cache, and a short
    file (error 84) caches only the frames it has."""
    wd = tmp_path / "cache"; wd.mkdir()
    _write_small_case(wd)
    env = {"PREALPHA_B200_CACHE": "1"}
    rc, out = _run_general_inp(wd, env)
    assert rc == 1 and "wrote the binary frame cache" in out and "(12 steps)" in out          # (1 = warning 55 of xyz trajectories)
    cache = wd / "traj.xyz.pab200cache"
    assert cache.exists()
    blob = cache.read_bytes()
    assert blob[:8] == b"PAB200C1"
    # payload: [frame][atom][3] float32 in file order, bit-identical to strtof of the text
    data = np.frombuffer(blob[-12 * 24 * 12:], dtype=np.float32).reshape(12, 24, 3)
    text = [ln.split() for ln in open(wd / "traj.xyz") if len(ln.split()) == 4]
    exp = np.array([[np.float32(v) for v in t[1:]] for t in text], dtype=np.float32).reshape(12, 24, 3)
    assert np.array_equal(data, exp)
    rc, out = _run_general_inp(wd, env)
    assert rc == 1 and "reading the binary frame cache" in out and "wrote the binary frame cache" not in out
    assert "reading the trajectory...done (12 steps of 24 atoms" in out
    # without the variable the cache is ignored
    rc, out = _run_general_inp(wd)
    assert "binary frame cache" not in out
    # another molecule table: not this cache
    (wd / "molecular.inp").write_text(" 12\n 1\n +1 2 12\nquit\n")
    rc, out = _run_general_inp(wd, env)
    assert "reading the binary frame cache" not in out and "wrote the binary frame cache" in out
    # the source changes (size / mtime): re-parsed and re-written
    _write_small_case(wd, nsteps_declared=12, nsteps_written=13)
    rc, out = _run_general_inp(wd, env)
    assert "reading the binary frame cache" not in out and "wrote the binary frame cache" in out
    # a cache directory instead of "next to the trajectory"
    cdir = tmp_path / "elsewhere"; cdir.mkdir()
    rc, out = _run_general_inp(wd, {"PREALPHA_B200_CACHE": str(cdir)})
    assert (cdir / "traj.xyz.pab200cache").exists()


@pytest.mark.parametrize("fmt,block,trailing_newline", [("xyz", 64, True), ("xyz", 257, False), ("xyz", 5000, True), ("lmp", 97, True),
                                                        ("lmp", 100000, False), ("xyz", 1 << 20, True)])
def test_block_reader_is_independent_of_block_and_thread_boundaries(fmt, block, trailing_newline, tmp_path):
    """The threaded text reader (read_frames: blocks read by several threads, newlines counted on byte ranges, frames cut
    at the newline count, lines parsed on all host threads) with blocks far smaller than a frame, blocks holding many
    frames, a last line without newline, lmp headers and a two-type molecule table: the float32 payload it leaves in the
    binary frame cache equals numpy's correctly rounded float32 of every number in the text, whatever the block size
    (PA_READ_BLOCK is the test hook for the bytes per read)."""
    wd = tmp_path / "blk"; wd.mkdir()
    rng = np.random.default_rng(block)
    nsteps, n1, n2 = 13, 7, 5                                     # type 1: 7 molecules of 2 atoms, type 2: 5 atoms
    n = 2 * n1 + n2
    lines, exp = [], np.zeros((nsteps, n, 3), dtype=np.float32)
    for s in range(nsteps):
        if fmt == "xyz":
            lines += [f"{n}", f"step {s}"]
        else:
            lines += ["ITEM: TIMESTEP", f"{s * 10}", "ITEM: NUMBER OF ATOMS", f"{n}", "ITEM: BOX BOUNDS pp pp pp", "0.0 12.0", "0.0 12.0", "0.0 12.0",
                      "ITEM: ATOMS element xu yu zu"]
        for a in range(n):
            txt = [f"{v:.{int(rng.integers(1, 9))}f}" if rng.random() < 0.8 else f"{v:.4e}" for v in (rng.random(3) - 0.2) * 15.0]
            pad = " " * int(rng.integers(1, 4))
            lines.append(("  " if a % 3 == 0 else "") + ("C" if a < 2 * n1 else "Na") + pad + pad.join(txt) + ("  " if a % 5 == 0 else ""))
            exp[s, a] = [np.float32(t) for t in txt]
    (wd / f"traj.{fmt}").write_text("\n".join(lines) + ("\n" if trailing_newline else ""))
    (wd / "molecular.inp").write_text(f" {nsteps}\n 2\n +0 2 {n1}\n +0 1 {n2}\nquit\n")
    (wd / "general.inp").write_text(f' "traj.{fmt}"\n "molecular.inp"\n "./"\n "./"\n "./"\n' + (" cubic_box_edge 0.0 12.0\n" if fmt == "xyz" else "") +
                                    f" trajectory_type {fmt}\n sequential_read F\n quit\n")
    rc, out = _run_general_inp(wd, {"PREALPHA_B200_CACHE": "1", "PA_READ_BLOCK": str(block)})
    assert rc == (1 if fmt == "xyz" else 0) and f"reading the trajectory...done ({nsteps} steps of {n} atoms" in out, out
    blob = (wd / f"traj.{fmt}.pab200cache").read_bytes()
    data = np.frombuffer(blob[-12 * n * nsteps:], dtype=np.float32).reshape(nsteps, n, 3)
    assert np.array_equal(data.view(np.uint32), exp.view(np.uint32))
