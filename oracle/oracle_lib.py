"""TEST INFRASTRUCTURE: ctypes loader for the CPU restatement (oracle/libpa_oracle.so)
and runner for the reference's own shipped binary (oracle/_ref/prealpha).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under prealpha_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import time

import numpy as np

from prealpha_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libpa_oracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "prealpha")

_I64P = C.POINTER(C.c_int64)
_SIG = {
    "orc_abi_version": (C.c_int, []),
    "orc_job_setup": (C.c_int, [C.POINTER(capi.pa_traj), C.POINTER(capi.pa_job_request), C.POINTER(capi.pa_job)]),
    "orc_distribution_histogram": (C.c_int, [C.POINTER(capi.pa_traj), C.POINTER(capi.pa_job), C.c_int,
                                             _I64P, _I64P, _I64P]),
    "orc_distribution_histogram_shard": (C.c_int, [C.POINTER(capi.pa_traj), C.POINTER(capi.pa_job), C.c_int, C.c_int,
                                                   C.c_int, _I64P, _I64P, _I64P]),
    "orc_transfer_histogram": (C.c_int, [C.POINTER(capi.pa_traj), C.POINTER(capi.pa_job), _I64P, C.c_int64,
                                         C.c_int64, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "orc_write_distribution": (C.c_int, [C.POINTER(capi.pa_traj), C.POINTER(capi.pa_job), C.c_int, C.c_char_p,
                                         C.c_char_p, C.POINTER(C.c_float), C.c_float]),
    "orc_format_real4": (C.c_int, [C.c_float, C.c_char_p]),
    "orc_distance_subjobs": (C.c_int, [C.POINTER(capi.pa_traj), C.POINTER(capi.pa_dist_job), C.POINTER(capi.pa_dist_result)]),
    "orc_contact_distance": (C.c_int, [C.POINTER(capi.pa_traj), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "orc_neighbour_pairs": (C.c_int, [C.POINTER(capi.pa_traj), C.POINTER(capi.pa_neigh_request), C.POINTER(capi.pa_neigh_pair),
                                      C.c_int64, _I64P]),
    "orc_wrap_full": (C.c_int, [C.POINTER(capi.pa_traj)]),
    "orc_unwrap_full": (C.c_int, [C.POINTER(capi.pa_traj)]),
    "orc_read_xyz": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_char_p]),
    "orc_read_traj": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_char_p]),
}
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libpa_oracle.so"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _lib = C.CDLL(ORACLE_SO)
        for name, (res, args) in _SIG.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def job_setup(traj: capi.Trajectory, req: capi.pa_job_request) -> capi.pa_job:
    job = capi.pa_job()
    rc = load().orc_job_setup(traj.ptr, C.byref(req), C.byref(job))
    if rc != 0:
        raise RuntimeError(f"orc_job_setup -> {rc}")
    return job


def histogram(traj: capi.Trajectory, job: capi.pa_job, nthreads=0, shard=(0, 1)):
    hist = np.zeros(job.bin_a * job.bin_b, dtype=np.int64)
    w, o = C.c_int64(0), C.c_int64(0)
    rc = load().orc_distribution_histogram_shard(traj.ptr, C.byref(job), shard[0], shard[1], nthreads,
                                                 hist.ctypes.data_as(_I64P), C.byref(w), C.byref(o))
    assert rc == 0
    return hist, w.value, o.value


def histogram_frame(traj: capi.Trajectory, job: capi.pa_job, step=0, nthreads=0):
    """One frame, OpenMP over blocks of reference molecules (same loop body and integers as `histogram`;
    for frames of the benchmark's size, where the frame loop offers no parallelism)."""
    hist = np.zeros(job.bin_a * job.bin_b, dtype=np.int64)
    w, o = C.c_int64(0), C.c_int64(0)
    rc = load().orc_distribution_histogram_frame(traj.ptr, C.byref(job), step, nthreads,
                                                 hist.ctypes.data_as(_I64P), C.byref(w), C.byref(o))
    assert rc == 0
    return hist, w.value, o.value


def transfer(traj, job, hist, within, oob, verbose=1):
    g = np.zeros(job.bin_a * job.bin_b, dtype=np.float32)
    rho = C.c_float(0)
    hist = np.ascontiguousarray(hist, dtype=np.int64)
    rc = load().orc_transfer_histogram(traj.ptr, C.byref(job), hist.ctypes.data_as(_I64P), within, oob, verbose,
                                       g.ctypes.data_as(C.POINTER(C.c_float)), C.byref(rho))
    return g, rho.value, rc


def write(traj, job, refnum, prefix, input_name, g, rho):
    g = np.ascontiguousarray(g, dtype=np.float32).copy()
    rc = load().orc_write_distribution(traj.ptr, C.byref(job), refnum, prefix.encode(), input_name.encode(),
                                       g.ctypes.data_as(C.POINTER(C.c_float)), rho)
    assert rc == 0, rc
    return g


def wrap_full(traj):
    """In-place wrap_full (Molecular:1703): rewrites the float32 coordinates held by `traj`."""
    assert load().orc_wrap_full(traj.ptr) == 0


def unwrap_full(traj):
    """In-place wrap_full + unwrap_full_SEQUENTIAL twice (Molecular:4806-4816, developers version)."""
    wrap_full(traj)
    assert load().orc_unwrap_full(traj.ptr) == 0
    assert load().orc_unwrap_full(traj.ptr) == 0


def distance_subjobs(traj, job):
    res = (capi.pa_dist_result * job.nsubjobs)()
    rc = load().orc_distance_subjobs(traj.ptr, C.byref(job), res)
    assert rc == 0
    return [r.as_dict() for r in res]


def contact_distance(traj, timestep, type_, nthreads=0):
    """(largest intramolecular, smallest intramolecular, smallest intermolecular) as float32 (Debug:586-623)."""
    out = (C.c_float * 3)()
    rc = load().orc_contact_distance(traj.ptr, timestep, type_, nthreads or (os.cpu_count() or 1), out)
    assert rc == 0, rc
    return tuple(np.float32(v) for v in out)


def neighbour_pairs(traj, rq, capacity=1 << 22):
    out = np.zeros((capacity, 2), dtype=np.int32)
    count = C.c_int64(0)
    rc = load().orc_neighbour_pairs(traj.ptr, C.byref(rq), out.ctypes.data_as(C.POINTER(capi.pa_neigh_pair)), capacity, C.byref(count))
    assert rc == 0 and count.value <= capacity, (rc, count.value)
    return out[:count.value]


def format_real4(x) -> str:
    buf = C.create_string_buffer(64)
    load().orc_format_real4(float(np.float32(x)), buf)
    return buf.value.decode()


def read_xyz(path, natoms_total, nsteps, header_lines=2):
    out = np.zeros((nsteps, natoms_total, 3), dtype=np.float32)
    el = C.create_string_buffer(natoms_total * 4)
    rc = load().orc_read_traj(path.encode(), natoms_total, nsteps, header_lines, out.ctypes.data_as(C.POINTER(C.c_float)), el)
    if rc != 0:
        raise RuntimeError(f"orc_read_xyz({path}) -> {rc}")
    elements = [el.raw[i * 4:i * 4 + 4].split(b"\0")[0].decode() for i in range(natoms_total)]
    return out, elements


# --------------------------------------------------------------------------
# the reference's own binary

def ref_available() -> bool:
    return os.path.exists(REF_BIN) and os.path.exists(os.path.join(REF_DIR, "lib", "libgfortran.so.5"))


def _ref_env():
    env = dict(os.environ)
    sc = None
    p = os.path.join(REF_DIR, "scipy_libs_path")
    if os.path.exists(p):
        sc = open(p).read().strip()
    if not sc or not os.path.isdir(sc):
        import scipy
        sc = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    env["LD_LIBRARY_PATH"] = os.path.join(REF_DIR, "lib") + ":" + sc + ":" + env.get("LD_LIBRARY_PATH", "")
    return env


def run_reference(workdir, general="general.inp", timeout=3600):
    """Run the shipped prealpha binary in workdir; returns (stdout, wall_seconds)."""
    t0 = time.time()
    p = subprocess.run([REF_BIN, general], cwd=workdir, env=_ref_env(), capture_output=True, timeout=timeout)
    out = p.stdout.decode(errors="replace")
    if p.returncode != 0:
        raise RuntimeError(f"reference binary failed ({p.returncode}): {out[-2000:]} {p.stderr.decode()[-2000:]}")
    return out, time.time() - t0


def ref_probe() -> bool:
    """True when the staged binary actually starts on this machine."""
    if not ref_available():
        return False
    try:
        p = subprocess.run([REF_BIN, "/nonexistent_general.inp"], env=_ref_env(), capture_output=True, timeout=60,
                           stdin=subprocess.DEVNULL)
        return b"libgfortran" not in p.stderr and p.returncode in (0, 1, 2)
    except Exception:
        return False


_WITHIN = re.compile(r"within bin (-?\d+), out of bounds (-?\d+)\.")
_TOOK = re.compile(r"End of parallelised section, took\s+([0-9.]+)\s+(\w+)")


def parse_within(stdout):
    return [(int(a), int(b)) for a, b in _WITHIN.findall(stdout)]


def parse_took_seconds(stdout):
    unit = {"seconds": 1.0, "milliseconds": 1e-3, "microseconds": 1e-6, "minutes": 60.0, "hours": 3600.0}
    return [float(v) * unit.get(u, 1.0) for v, u in _TOOK.findall(stdout)]


def write_xyz(path, elements, frames, fmt="%.5f"):
    """frames: [nsteps, natoms, 3]; elements: list of natoms strings."""
    with open(path, "w") as f:
        n = len(elements)
        for s in range(frames.shape[0]):
            f.write(f"{n}\nstep {s}\n")
            fr = frames[s]
            f.write("".join(f"{elements[a]} {fmt % fr[a, 0]} {fmt % fr[a, 1]} {fmt % fr[a, 2]}\n" for a in range(n)))


def write_lmp(path, elements, frames, box_lo, box_hi, fmt="%.6f"):
    """LAMMPS dump with the 9 header lines the reference expects (Molecular:4833-4871)."""
    with open(path, "w") as f:
        n = len(elements)
        for s in range(frames.shape[0]):
            f.write(f"ITEM: TIMESTEP\n{s * 1000}\nITEM: NUMBER OF ATOMS\n{n}\nITEM: BOX BOUNDS pp pp pp\n")
            for k in range(3):
                f.write(f"{float(box_lo[k])!r} {float(box_hi[k])!r}\n")
            f.write("ITEM: ATOMS element xu yu zu\n")
            fr = frames[s]
            f.write("".join(f"{elements[a]} {fmt % fr[a, 0]} {fmt % fr[a, 1]} {fmt % fr[a, 2]}\n" for a in range(n)))


def write_case(workdir, traj_name, types, nsteps, box, dist_inputs, threads=8, extra_general=()):
    """Write general.inp / molecular.inp / distribution inputs for a reference run.
    types: list of (charge, natoms, nmol); dist_inputs: {filename: text}."""
    os.makedirs(os.path.join(workdir, "out"), exist_ok=True)
    with open(os.path.join(workdir, "molecular.inp"), "w") as f:
        f.write(f" {nsteps}\n {len(types)}\n")
        for ch, na, nm in types:
            f.write(f" {ch:+d} {na} {nm}\n")
        f.write("quit\n")
    with open(os.path.join(workdir, "general.inp"), "w") as f:
        f.write(f' "{traj_name}"\n "molecular.inp"\n "./"\n "./"\n "./out/"\n')
        if box is not None:
            f.write(f" cubic_box_edge {box[0]} {box[1]}\n")
        f.write(f" sequential_read F\n parallel_operation T\n set_threads {threads}\n")
        for line in extra_general:
            f.write(f" {line}\n")
        for name in dist_inputs:
            f.write(f' distribution "{name}"\n')
        f.write(" quit\n")
    for name, text in dist_inputs.items():
        with open(os.path.join(workdir, name), "w") as f:
            f.write(text)
