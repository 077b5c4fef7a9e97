"""Thin Python face of the C ABI (ctypes).  Mirrors what the reference's
perform_distribution_analysis (Distribution:1187-1228) does per reference line:
set-up -> histogram (GPU) -> transfer_histogram -> write_distribution_functions."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from . import capi

_I64P = C.POINTER(C.c_int64)


def device_count() -> int:
    return capi.load().pa_device_count()


def job_setup(traj: capi.Trajectory, req: capi.pa_job_request) -> capi.pa_job:
    job = capi.pa_job()
    capi.check(capi.load().pa_job_setup(traj.ptr, C.byref(req), C.byref(job)), "pa_job_setup")
    return job


def fp64_issue_rate(device=0) -> float:
    """Measured FP64-pipe instruction issue rate (DSUB/DMUL chains, no FMA credit) of this GPU."""
    lib = capi.load()
    lib.pa_debug_fp64_issue_rate.restype = C.c_int
    lib.pa_debug_fp64_issue_rate.argtypes = [C.c_int, C.POINTER(C.c_double)]
    v = C.c_double(0)
    capi.check(lib.pa_debug_fp64_issue_rate(device, C.byref(v)), "pa_debug_fp64_issue_rate")
    return v.value


def smem_atomic_rate(device=0) -> float:
    """Shared-memory atomic increments per second with the pair kernels' access pattern (32 different counters per
    warp instruction), measured live on `device`: the roofline denominator of the cell-sorted RDF kernels."""
    lib = capi.load()
    lib.pa_debug_smem_atomic_rate.restype = C.c_int
    lib.pa_debug_smem_atomic_rate.argtypes = [C.c_int, C.POINTER(C.c_double)]
    v = C.c_double(0)
    capi.check(lib.pa_debug_smem_atomic_rate(device, C.byref(v)), "pa_debug_smem_atomic_rate")
    return v.value


def make_exec(device=0, shard_rank=0, shard_count=1, frames_per_batch=0, flags=0, tile_shard_rank=0, tile_shard_count=1,
              ndevices=0) -> capi.pa_exec:
    ex = capi.pa_exec()
    ex.device, ex.shard_rank, ex.shard_count, ex.frames_per_batch, ex.flags = device, shard_rank, shard_count, frames_per_batch, flags
    ex.tile_shard_rank, ex.tile_shard_count, ex.ndevices = tile_shard_rank, tile_shard_count, ndevices
    return ex


def _jobs_array(jobs: Sequence[capi.pa_job]):
    arr = (capi.pa_job * len(jobs))()
    for i, j in enumerate(jobs):
        C.memmove(C.byref(arr[i]), C.byref(j), C.sizeof(capi.pa_job))
    return arr


def histogram_multi(traj: capi.Trajectory, jobs: Sequence[capi.pa_job], exec_: capi.pa_exec | None = None):
    """Host buffers in, int64 histograms out (the end-to-end path).  Returns (hists, within, oob, stats)."""
    lib = capi.load()
    n = len(jobs)
    arr = _jobs_array(jobs)
    hists = [np.zeros(j.bin_a * j.bin_b, dtype=np.int64) for j in jobs]
    hp = (_I64P * n)(*[h.ctypes.data_as(_I64P) for h in hists])
    within = np.zeros(n, dtype=np.int64)
    oob = np.zeros(n, dtype=np.int64)
    st = capi.pa_stats()
    ex = exec_ if exec_ is not None else make_exec()
    rc = lib.pa_distribution_histogram_multi(traj.ptr, arr, n, C.byref(ex), hp, within.ctypes.data_as(_I64P),
                                             oob.ctypes.data_as(_I64P), C.byref(st))
    capi.check(rc, "pa_distribution_histogram_multi")
    return hists, within, oob, st


def histogram(traj: capi.Trajectory, job: capi.pa_job):
    lib = capi.load()
    hist = np.zeros(job.bin_a * job.bin_b, dtype=np.int64)
    w, o = C.c_int64(0), C.c_int64(0)
    capi.check(lib.pa_distribution_histogram(traj.ptr, C.byref(job), hist.ctypes.data_as(_I64P), C.byref(w), C.byref(o)),
               "pa_distribution_histogram")
    return hist, w.value, o.value


def transfer(traj, job, hist, within, oob, verbose=1):
    g = np.zeros(job.bin_a * job.bin_b, dtype=np.float32)
    rho = C.c_float(0)
    hist = np.ascontiguousarray(hist, dtype=np.int64)
    rc = capi.load().pa_transfer_histogram(traj.ptr, C.byref(job), hist.ctypes.data_as(_I64P), int(within), int(oob),
                                           verbose, g.ctypes.data_as(C.POINTER(C.c_float)), C.byref(rho))
    return g, rho.value, rc


def write(traj, job, refnum, prefix, input_name, g, rho):
    g = np.ascontiguousarray(g, dtype=np.float32).copy()
    capi.check(capi.load().pa_write_distribution(traj.ptr, C.byref(job), refnum, prefix.encode(), input_name.encode(),
                                                 g.ctypes.data_as(C.POINTER(C.c_float)), rho), "pa_write_distribution")
    return g


def distance_subjobs(traj, job, exec_=None):
    """Row X through the C ABI: list of result dicts, one per subjob."""
    res = (capi.pa_dist_result * job.nsubjobs)()
    ex = exec_ if exec_ is not None else make_exec()
    capi.check(capi.load().pa_distance_subjobs(traj.ptr, C.byref(job), C.byref(ex), res), "pa_distance_subjobs")
    return [r.as_dict() for r in res]


def contact_distance(traj, timestep, type_, exec_=None):
    """contact_distance (Debug:586-623) for one molecule type at one time step (both 1-based):
    (largest intramolecular, smallest intramolecular, smallest intermolecular) as float32, plus the pair count."""
    res = capi.pa_contact_result()
    ex = exec_ if exec_ is not None else make_exec()
    capi.check(capi.load().pa_contact_distance(traj.ptr, timestep, type_, C.byref(ex), C.byref(res)), "pa_contact_distance")
    return (np.float32(res.largest_intramolecular), np.float32(res.smallest_intramolecular),
            np.float32(res.smallest_intermolecular)), int(res.pairs_evaluated)


def neighbour_pairs(traj, rq, capacity=1 << 20, exec_=None):
    """pa_neighbour_pairs: int32 array [count][2] of (ia, ib), sorted in the reference's loop order, and the number of tests."""
    ex = exec_ if exec_ is not None else make_exec()
    while True:
        out = np.zeros((max(capacity, 1), 2), dtype=np.int32)
        count, tests = C.c_int64(0), C.c_int64(0)
        capi.check(capi.load().pa_neighbour_pairs(traj.ptr, C.byref(rq), C.byref(ex), out.ctypes.data_as(C.POINTER(capi.pa_neigh_pair)),
                                                  capacity, C.byref(count), C.byref(tests)), "pa_neighbour_pairs")
        if count.value <= capacity:
            return out[:count.value], tests.value
        capacity = int(count.value)


def format_real4(x) -> str:
    buf = C.create_string_buffer(64)
    capi.load().pa_format_real4(float(np.float32(x)), buf)
    return buf.value.decode()


class Session:
    """Device-resident frames + accumulators (pa_session_*): upload once, run many times."""

    def __init__(self, traj: capi.Trajectory, jobs: Sequence[capi.pa_job], max_frames: int, exec_: capi.pa_exec | None = None):
        self.lib = capi.load()
        self.jobs = list(jobs)
        self._arr = _jobs_array(jobs)
        self._ex = exec_ if exec_ is not None else make_exec()
        self.h = C.c_void_p()
        capi.check(self.lib.pa_session_create(traj.ptr, self._arr, len(jobs), C.byref(self._ex), max_frames, C.byref(self.h)),
                   "pa_session_create")

    def upload(self, traj, first_step, nframes, stride=1):
        capi.check(self.lib.pa_session_upload(self.h, traj.ptr, first_step, nframes, stride), "pa_session_upload")

    def run(self, first=None, count=None):
        if first is None:
            capi.check(self.lib.pa_session_run(self.h), "pa_session_run")
        else:
            capi.check(self.lib.pa_session_run_frames(self.h, first, 1 if count is None else count), "pa_session_run_frames")

    def download(self, reset=False):
        n = len(self.jobs)
        hists = [np.zeros(j.bin_a * j.bin_b, dtype=np.int64) for j in self.jobs]
        hp = (_I64P * n)(*[h.ctypes.data_as(_I64P) for h in hists])
        within = np.zeros(n, dtype=np.int64)
        oob = np.zeros(n, dtype=np.int64)
        capi.check(self.lib.pa_session_download(self.h, hp, within.ctypes.data_as(_I64P), oob.ctypes.data_as(_I64P), int(reset)),
                   "pa_session_download")
        return hists, within, oob

    def stats(self) -> capi.pa_stats:
        st = capi.pa_stats()
        capi.check(self.lib.pa_session_stats(self.h, C.byref(st)), "pa_session_stats")
        return st

    def device_accumulators(self):
        p, n = C.c_void_p(), C.c_int64(0)
        capi.check(self.lib.pa_session_device_accumulators(self.h, C.byref(p), C.byref(n)), "pa_session_device_accumulators")
        return p.value, n.value

    def close(self):
        if self.h:
            self.lib.pa_session_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
