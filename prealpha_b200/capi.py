"""ctypes mirror of include/prealpha_b200.h.

This is plumbing for the Python callers (tests, bench.py): the product is the
C-ABI shared library ``prealpha_b200/libprealpha_b200.so`` built from
``prealpha_b200/csrc``.  There is no Python or CPU fallback: loading fails loudly
when the library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprealpha_b200.so")

PA_MODE_CDF = 0
PA_MODE_PDF = 1
PA_MODE_CHARGE_ARM = 2
PA_ERR_NO_DEVICE = -1
PA_ERR_CUDA = -2
PA_ERR_BAD_ARG = -3
PA_ERR_UNSUPPORTED = -4
PA_ERR_IO = -5
PA_EXEC_FORCE_GENERAL = 1
PA_EXEC_NO_CELLS = 2
PA_EXEC_NO_SYMMETRY = 4
PA_EXEC_NO_F32 = 8


class pa_type(C.Structure):
    _fields_ = [
        ("charge", C.c_int32), ("natoms", C.c_int32), ("nmol", C.c_int32), ("_pad0", C.c_int32),
        ("mass", C.c_double), ("realcharge", C.c_float), ("_pad1", C.c_int32),
        ("atom_mass", C.POINTER(C.c_double)), ("atom_charge", C.POINTER(C.c_double)),
        ("xyz", C.POINTER(C.c_float)),
    ]


class pa_traj(C.Structure):
    _fields_ = [
        ("ntypes", C.c_int32), ("nsteps", C.c_int32), ("types", C.POINTER(pa_type)),
        ("box_lo", C.c_double * 3), ("box_hi", C.c_double * 3), ("max_dist_sq", C.c_double),
        ("com_boost", C.c_int32), ("wrap_trajectory", C.c_int32),
    ]


class pa_job_request(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("ref_xyz", C.c_int32 * 3), ("ref_type", C.c_int32), ("obs_type", C.c_int32),
        ("bin_a", C.c_int32), ("bin_b", C.c_int32), ("sampling_interval", C.c_int32),
        ("weigh_charge", C.c_int32), ("use_coc", C.c_int32), ("subtract_uniform", C.c_int32),
        ("maxdist_optimize", C.c_int32), ("maxdist", C.c_float), ("normalise_clm", C.c_int32),
    ]


class pa_job(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("ref_type", C.c_int32), ("obs_type", C.c_int32), ("ref_xyz", C.c_int32 * 3),
        ("ref_vec", C.c_double * 3), ("rot", C.c_double * 9),
        ("maxdist", C.c_float), ("step_a", C.c_float), ("step_b", C.c_float), ("shift", C.c_float),
        ("bin_a", C.c_int32), ("bin_b", C.c_int32), ("sampling_interval", C.c_int32),
        ("weigh_charge", C.c_int32), ("use_coc", C.c_int32), ("subtract_uniform", C.c_int32),
        ("aligned", C.c_int32), ("normalise_clm", C.c_int32),
    ]


class pa_exec(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
        ("frames_per_batch", C.c_int32), ("flags", C.c_int32),
        ("tile_shard_rank", C.c_int32), ("tile_shard_count", C.c_int32), ("ndevices", C.c_int32),
    ]


class pa_stats(C.Structure):
    _fields_ = [
        ("pairs_evaluated", C.c_int64), ("pairs_binned", C.c_int64), ("frames_done", C.c_int64),
        ("kernel_launches", C.c_int64), ("exception_pairs", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("kernel_ms", C.c_double), ("total_ms", C.c_double),
        ("cell_pair_evals", C.c_int64), ("smem_atomics", C.c_int64), ("unique_pairs_counted", C.c_int64),
        ("dominant_launches", C.c_int64), ("dominant_ms", C.c_double),
        ("dominant_pair_evals", C.c_int64), ("dominant_atomics", C.c_int64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class pa_dist_subjob(C.Structure):
    _fields_ = [("n_ref", C.c_int32), ("n_obs", C.c_int32), ("ref", C.POINTER(C.c_int32)), ("obs", C.POINTER(C.c_int32)),
                ("weight", C.c_int64)]


class pa_dist_job(C.Structure):
    _fields_ = [
        ("intermolecular", C.c_int32), ("nsubjobs", C.c_int32), ("subjobs", C.POINTER(pa_dist_subjob)),
        ("nsteps", C.c_int32), ("sampling_interval", C.c_int32),
        ("calc_exp", C.c_int32), ("calc_ffc", C.c_int32), ("calc_stdev", C.c_int32),
        ("maxdist", C.c_float), ("exponent", C.c_float),
    ]


class pa_dist_result(C.Structure):
    _fields_ = [
        ("closest_average", C.c_double), ("closest_stdev", C.c_double),
        ("exponential_average", C.c_double), ("exponential_stdev", C.c_double), ("exponential_weight", C.c_double),
        ("ffc_average", C.c_double), ("ffcexp_average", C.c_double), ("ffc_weight", C.c_double),
        ("missed_neighbours", C.c_int64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class pa_contact_result(C.Structure):
    _fields_ = [
        ("largest_intramolecular", C.c_float), ("smallest_intramolecular", C.c_float), ("smallest_intermolecular", C.c_float),
        ("pad_", C.c_int32), ("pairs_evaluated", C.c_int64),
    ]


class pa_neigh_request(C.Structure):
    _fields_ = [
        ("timestep", C.c_int32), ("n_a", C.c_int32), ("n_b", C.c_int32),
        ("a", C.POINTER(C.c_int32)), ("b", C.POINTER(C.c_int32)),
        ("a_class", C.POINTER(C.c_int32)), ("b_class", C.POINTER(C.c_int32)),
        ("n_class_a", C.c_int32), ("n_class_b", C.c_int32),
        ("cutoff_sq", C.POINTER(C.c_float)),
        ("same_list", C.c_int32), ("skip_same_molecule", C.c_int32),
    ]


class pa_neigh_pair(C.Structure):
    _fields_ = [("ia", C.c_int32), ("ib", C.c_int32)]


def make_neigh_request(timestep, a, b, cutoff_sq, a_class=None, b_class=None, same_list=False, skip_same_molecule=False):
    """a, b: int arrays [n][3] of (molecule type, molecule index, atom index), 1-based; cutoff_sq: float32 scalar or matrix.
    Returns (request, keepalive)."""
    a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 3)
    b = a if same_list else np.ascontiguousarray(b, dtype=np.int32).reshape(-1, 3)
    cut = np.ascontiguousarray(np.atleast_2d(np.asarray(cutoff_sq, dtype=np.float32)))
    keep = [a, b, cut]
    rq = pa_neigh_request()
    rq.timestep, rq.n_a, rq.n_b = int(timestep), a.shape[0], b.shape[0]
    rq.a = a.ctypes.data_as(C.POINTER(C.c_int32)); rq.b = b.ctypes.data_as(C.POINTER(C.c_int32))
    for name, arr in (("a_class", a_class), ("b_class", b_class)):
        if arr is not None:
            arr = np.ascontiguousarray(arr, dtype=np.int32); keep.append(arr)
            setattr(rq, name, arr.ctypes.data_as(C.POINTER(C.c_int32)))
    rq.n_class_a, rq.n_class_b = cut.shape
    rq.cutoff_sq = cut.ctypes.data_as(C.POINTER(C.c_float))
    rq.same_list, rq.skip_same_molecule = int(bool(same_list)), int(bool(skip_same_molecule))
    return rq, keep


# symbol -> (restype, argtypes); every entry of include/prealpha_b200.h
_I64P = C.POINTER(C.c_int64)
SIGNATURES = {
    "pa_abi_version": (C.c_int, []),
    "pa_last_error": (C.c_char_p, []),
    "pa_device_count": (C.c_int, []),
    "pa_trim": (None, []),
    "pa_job_setup": (C.c_int, [C.POINTER(pa_traj), C.POINTER(pa_job_request), C.POINTER(pa_job)]),
    "pa_distribution_histogram": (C.c_int, [C.POINTER(pa_traj), C.POINTER(pa_job), _I64P, _I64P, _I64P]),
    "pa_distribution_histogram_multi": (C.c_int, [C.POINTER(pa_traj), C.POINTER(pa_job), C.c_int32,
                                                  C.POINTER(pa_exec), C.POINTER(_I64P), _I64P, _I64P,
                                                  C.POINTER(pa_stats)]),
    "pa_session_create": (C.c_int, [C.POINTER(pa_traj), C.POINTER(pa_job), C.c_int32, C.POINTER(pa_exec),
                                    C.c_int32, C.POINTER(C.c_void_p)]),
    "pa_session_upload": (C.c_int, [C.c_void_p, C.POINTER(pa_traj), C.c_int32, C.c_int32, C.c_int32]),
    "pa_session_run": (C.c_int, [C.c_void_p]),
    "pa_session_run_frames": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "pa_session_download": (C.c_int, [C.c_void_p, C.POINTER(_I64P), _I64P, _I64P, C.c_int32]),
    "pa_session_stats": (C.c_int, [C.c_void_p, C.POINTER(pa_stats)]),
    "pa_session_device_accumulators": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), _I64P]),
    "pa_session_destroy": (None, [C.c_void_p]),
    "pa_transfer_histogram": (C.c_int, [C.POINTER(pa_traj), C.POINTER(pa_job), _I64P, C.c_int64, C.c_int64,
                                        C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "pa_write_distribution": (C.c_int, [C.POINTER(pa_traj), C.POINTER(pa_job), C.c_int32, C.c_char_p,
                                        C.c_char_p, C.POINTER(C.c_float), C.c_float]),
    "pa_format_real4": (C.c_int, [C.c_float, C.c_char_p]),
    "pa_distance_subjobs": (C.c_int, [C.POINTER(pa_traj), C.POINTER(pa_dist_job), C.POINTER(pa_exec), C.POINTER(pa_dist_result)]),
    "pa_contact_distance": (C.c_int, [C.POINTER(pa_traj), C.c_int32, C.c_int32, C.POINTER(pa_exec), C.POINTER(pa_contact_result)]),
    "pa_neighbour_pairs": (C.c_int, [C.POINTER(pa_traj), C.POINTER(pa_neigh_request), C.POINTER(pa_exec), C.POINTER(pa_neigh_pair),
                                     C.c_int64, _I64P, _I64P]),
    "pa_run_general_input": (C.c_int, [C.c_char_p, C.POINTER(pa_exec)]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def bind(lib, signatures=SIGNATURES, prefix_from="pa_", prefix_to="pa_"):
    for name, (res, args) in signatures.items():
        fn = getattr(lib, name.replace(prefix_from, prefix_to, 1) if prefix_from != prefix_to else name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """Load libprealpha_b200.so; raise (never fall back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for this path.")
        _lib = bind(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL))
        if _lib.pa_abi_version() != 2:
            raise RuntimeError("prealpha_b200 ABI mismatch")
    return _lib


def check(rc: int, what: str = "call"):
    if rc != 0:
        msg = load().pa_last_error()
        raise RuntimeError(f"prealpha_b200 {what} failed with code {rc}: {msg.decode() if msg else ''}")


# --------------------------------------------------------------------------
# helpers to build the C structs from numpy arrays

# Molecular:9-92 -- REAL(4) element masses (value widened to REAL(8) when stored)
ELEMENT_MASS = {
    "H": 1.008, "He": 4.002, "Li": 6.94, "Be": 9.012, "B": 10.81, "C": 12.011, "N": 14.007, "O": 15.999,
    "F": 18.998, "Ne": 20.1797, "Na": 22.989, "Mg": 24.305, "Al": 26.981, "Si": 28.085, "P": 30.973762,
    "S": 32.06, "Cl": 35.45, "Ar": 39.95, "K": 39.0983, "Ca": 40.078, "Sc": 44.955, "Ti": 47.867,
    "V": 50.9415, "Cr": 51.9961, "Mn": 54.938, "Fe": 55.845, "Co": 58.933, "Ni": 58.6934, "Cu": 63.546,
    "Zn": 65.38, "Ga": 69.723, "Ge": 72.63, "As": 74.921, "Se": 78.971, "Br": 79.904, "Kr": 83.798,
    "Rb": 85.4678, "Sr": 87.62, "Y": 88.905, "Zr": 91.222, "Nb": 92.906, "Mo": 95.95, "Ru": 101.07,
    "Rh": 102.905, "Pd": 106.42, "Ag": 107.8682, "Cd": 112.414, "In": 114.818, "Sn": 118.71, "Sb": 121.76,
    "Te": 127.6, "I": 126.904, "Xe": 131.293, "Cs": 132.905, "Ba": 137.327, "La": 138.905, "Ce": 140.116,
    "Pr": 140.907, "Nd": 144.242, "Sm": 150.36, "Eu": 151.964, "Gd": 157.249, "Tb": 158.925, "Dy": 162.5,
    "Ho": 164.93, "Er": 167.259, "Tm": 168.934, "Yb": 173.045, "Lu": 174.966, "Hf": 178.486, "Ta": 180.947,
    "W": 183.84, "Re": 186.207, "Os": 190.23, "Ir": 192.217, "Pt": 195.084, "Au": 196.966, "Hg": 200.592,
    "Tl": 204.38, "Pb": 207.2, "Bi": 208.98, "Th": 232.0377, "Pa": 231.035, "U": 238.028,
}


class Trajectory:
    """Owns the numpy buffers behind a ``pa_traj`` (keeps them alive)."""

    def __init__(self, types: Sequence[dict], box_lo, box_hi, *, xyz_format=True, wrap_trajectory=False,
                 max_dist_sq=None):
        """types: dicts with charge, natoms, nmol, elements (len natoms) or atom_mass, xyz
        (float32 array [nsteps, nmol, natoms, 3]); optional atom_charge."""
        self._keep = []
        n = len(types)
        self.types_arr = (pa_type * n)()
        nsteps = None
        for i, t in enumerate(types):
            xyz = np.ascontiguousarray(t["xyz"], dtype=np.float32)
            natoms, nmol = int(t["natoms"]), int(t["nmol"])
            assert xyz.shape[1:] == (nmol, natoms, 3), (xyz.shape, nmol, natoms)
            nsteps = xyz.shape[0] if nsteps is None else nsteps
            assert xyz.shape[0] == nsteps
            if "atom_mass" in t:
                am = np.asarray(t["atom_mass"], dtype=np.float64)
            else:
                am = np.array([np.float32(ELEMENT_MASS[e]) for e in t["elements"]], dtype=np.float32).astype(np.float64)
            # %mass: REAL8 accumulator += REAL4 element mass, file order (Molecular:4923)
            mass = 0.0
            for m in am:
                mass = mass + float(m)
            charge = int(t["charge"])
            if "atom_charge" in t:
                aq = np.asarray(t["atom_charge"], dtype=np.float64)
            elif natoms == 1:
                aq = np.array([float(charge)], dtype=np.float64)      # Molecular:4911-4913
            else:
                aq = np.zeros(natoms, dtype=np.float64)               # default element charges are 0.0
            am = np.ascontiguousarray(am); aq = np.ascontiguousarray(aq)
            self._keep += [xyz, am, aq]
            pt = self.types_arr[i]
            pt.charge, pt.natoms, pt.nmol = charge, natoms, nmol
            pt.mass = mass
            pt.realcharge = float(t.get("realcharge", float(charge)))
            pt.atom_mass = am.ctypes.data_as(C.POINTER(C.c_double))
            pt.atom_charge = aq.ctypes.data_as(C.POINTER(C.c_double))
            pt.xyz = xyz.ctypes.data_as(C.POINTER(C.c_float))
        self.xyz = [k for k in self._keep[0::3]]
        self.c = pa_traj()
        self.c.ntypes, self.c.nsteps = n, int(nsteps)
        self.c.types = C.cast(self.types_arr, C.POINTER(pa_type))
        lo = np.broadcast_to(np.asarray(box_lo, dtype=np.float64), (3,))
        hi = np.broadcast_to(np.asarray(box_hi, dtype=np.float64), (3,))
        for k in range(3):
            self.c.box_lo[k] = float(lo[k]); self.c.box_hi[k] = float(hi[k])
        size = hi - lo
        if max_dist_sq is None:
            # Molecular:938 / :4857, overwritten with 22500.0 by the xyz header reader (:4883)
            max_dist_sq = 22500.0 if xyz_format else float(size[1] ** 2 + np.sqrt(size[0] ** 2 + size[2] ** 2))
        self.c.max_dist_sq = max_dist_sq
        self.c.com_boost = int(all(int(t["natoms"]) == 1 for t in types))
        self.c.wrap_trajectory = int(bool(wrap_trajectory))

    @property
    def ptr(self):
        return C.byref(self.c)


def cubic_box_edge(lo: float, hi: float):
    """`cubic_box_edge lo hi` is read into REAL(4) variables (Main:108,228; Molecular:931-941)."""
    return float(np.float32(lo)), float(np.float32(hi))


def make_request(mode="pdf", vec=(0, 0, 1), ref_type=1, obs_type=1, bin_a=100, bin_b=100, sampling_interval=10,
                 weigh_charge=False, use_coc=False, subtract_uniform=False, maxdist_optimize=False,
                 maxdist=10.0, normalise_clm=False) -> pa_job_request:
    r = pa_job_request()
    r.mode = PA_MODE_CHARGE_ARM if mode in ("charge_arm", PA_MODE_CHARGE_ARM) else (PA_MODE_PDF if mode in ("pdf", "podf", PA_MODE_PDF) else PA_MODE_CDF)
    r.normalise_clm = int(normalise_clm)
    for k in range(3):
        r.ref_xyz[k] = int(vec[k])
    r.ref_type, r.obs_type = int(ref_type), int(obs_type)
    r.bin_a, r.bin_b, r.sampling_interval = int(bin_a), int(bin_b), int(sampling_interval)
    r.weigh_charge, r.use_coc, r.subtract_uniform = int(weigh_charge), int(use_coc), int(subtract_uniform)
    r.maxdist_optimize = int(maxdist_optimize)
    r.maxdist = float(np.float32(maxdist))
    return r
