"""Python mirror of the reference's distribution-input language (Distribution:328-616).

Used by the Python callers (tests, bench.py) to turn the text a prealpha user writes
into ``pa_job_request`` structs.  The C++ host (csrc/pa_host.cpp) has its own parser for the
file-level drop-in; both are checked against the reference binary's behaviour through the
golden fixtures.
"""
from __future__ import annotations

import shlex
from typing import List

from . import capi

DEFAULTS = dict(sampling_interval=10, bin_count_a=100, bin_count_b=100, maxdist=10.0, maxdist_optimize=False,
                subtract_uniform=False, weigh_charge=False, use_coc=False)   # Distribution:9-16


def _tokens(line: str) -> List[str]:
    """Fortran list-directed tokenisation: blanks/commas separate, quotes group, '/' terminates."""
    line = line.replace(",", " ")
    try:
        toks = shlex.split(line, posix=True)
    except ValueError:
        toks = line.split()
    out = []
    for t in toks:
        if t.startswith("/"):
            break
        out.append(t)
    return out


def _logical(tok: str) -> bool:
    t = tok.strip().lstrip(".").upper()
    if t[:1] == "T":
        return True
    if t[:1] == "F":
        return False
    raise ValueError(tok)


def parse_distribution_input(text: str, ntypes: int, box_limit: float | None = None) -> List[capi.pa_job_request]:
    """Returns one request per reference line.  Raises ValueError(code) on the reference's errors 99/33."""
    lines = text.splitlines()
    # list-directed READ skips blank lines
    lines = [l for l in lines if l.strip()]
    head = _tokens(lines[0])
    mode = head[0]
    mode = {"cydf": "cdf", "podf": "pdf"}.get(mode, mode)
    if mode not in ("cdf", "pdf", "charge_arm"):
        raise ValueError(99)
    nref = int(head[1])
    refs = []
    ncol = 4 if mode == "charge_arm" else 5                      # Distribution:409 / :434
    for k in range(nref):
        t = _tokens(lines[1 + k])
        v = [int(x) for x in t[:ncol]]
        if len(v) < ncol:
            raise ValueError(99)
        if ncol == 4:
            v.append(v[3])
        if v[3] > ntypes or v[3] < 1 or v[4] > ntypes:
            raise ValueError(33)
        refs.append(v)
    kw = dict(DEFAULTS)
    for line in lines[1 + nref:][:500]:
        t = _tokens(line)
        if not t:
            continue
        key = t[0]
        try:
            if key == "sampling_interval":
                kw["sampling_interval"] = int(t[1])
            elif key == "bin_count":
                kw["bin_count_a"] = kw["bin_count_b"] = int(t[1])
            elif key == "bin_count_a":
                kw["bin_count_a"] = int(t[1])
            elif key == "bin_count_b":
                kw["bin_count_b"] = int(t[1])
            elif key == "maxdist":
                kw["maxdist_optimize"] = False
                kw["maxdist"] = float(t[1])
            elif key in ("maxdist_optimize", "maxdist_optimise"):
                # the reference evaluates give_box_limit()/2.0 right here (Distribution:526-527);
                # pa_job_setup does the same REAL(4) arithmetic from the box in pa_traj.
                kw["maxdist_optimize"] = True
            elif key == "subtract_uniform":
                if mode in ("pdf", "charge_arm"):
                    kw["subtract_uniform"] = _logical(t[1])
            elif key == "weigh_charge":
                if mode != "charge_arm":
                    kw["weigh_charge"] = _logical(t[1])
            elif key in ("use_coc", "use_COC", "center_of_charge", "centre_of_charge"):
                if mode != "charge_arm":
                    kw["use_coc"] = _logical(t[1])
            elif key in ("normalise_CLM", "normalise_charge_arm", "normalize_CLM", "normalize_charge_arm", "normalise_clm"):
                if mode == "charge_arm":
                    kw["normalise_clm"] = _logical(t[1])
            elif key == "quit":
                break
        except (IndexError, ValueError):
            # error 100: keep/restore the default (Distribution:462-466 etc.)
            pass
        kw["bin_count_a"] = min(max(kw["bin_count_a"], 1), 1000)
        kw["bin_count_b"] = min(max(kw["bin_count_b"], 1), 1000)
    reqs = []
    for v in refs:
        reqs.append(capi.make_request(mode=mode, vec=v[:3], ref_type=v[3], obs_type=v[4], bin_a=kw["bin_count_a"],
                                      bin_b=kw["bin_count_b"], sampling_interval=kw["sampling_interval"],
                                      weigh_charge=kw["weigh_charge"], use_coc=kw["use_coc"],
                                      subtract_uniform=kw["subtract_uniform"],
                                      maxdist_optimize=kw["maxdist_optimize"], maxdist=kw["maxdist"],
                                      normalise_clm=kw.get("normalise_clm", False)))
    return reqs


# ---------------------------------------------------------------------------
# distance input (Distances:299-802): subjob lines with element wildcards, then keywords

def _is_int(tok):
    try:
        int(tok)
        return True
    except ValueError:
        return False


def parse_distance_input(text: str, type_elements, type_nmol, box_limit=None, sticky=None):
    """type_elements: per molecule type the list of element symbols; type_nmol: molecules per type.
    sticky: dict(calc_exp, calc_ffc) carried between invocations (Distances:288-296 does not reset them).
    Returns dict(inter, subjobs=[dict(ref, obs, weight, description, el1, el2)], nsteps, sampling_interval,
    calc_exp, calc_ffc, calc_stdev, maxdist, exponent)."""
    sticky = sticky if sticky is not None else {}
    lines = [l for l in text.splitlines() if l.strip()]
    head = _tokens(lines[0])
    mode = {"intramolecular": "intra", "intermolecular": "inter"}.get(head[0], head[0])
    if mode not in ("intra", "inter"):
        raise ValueError(135)
    inter = mode == "inter"
    nsub = int(head[1])
    ntypes = len(type_elements)

    def atoms_of(el, only_type=0):
        return [(t + 1, a + 1) for t in range(ntypes) if not only_type or t + 1 == only_type
                for a, e in enumerate(type_elements[t]) if e == el]

    subs = []
    for tk in (_tokens(l) for l in lines[1:1 + nsub]):
        wild, t1, a1, t2, a2, e1, e2 = None, 0, 0, 0, 0, "", ""
        if inter:
            if len(tk) >= 4 and all(_is_int(x) for x in tk[:4]):
                t1, a1, t2, a2 = (int(x) for x in tk[:4]); wild = 0
            elif len(tk) >= 3 and _is_int(tk[1]) and _is_int(tk[2]):
                e2, t1, a1, wild = tk[0][:2], int(tk[1]), int(tk[2]), 1          # swapped (notice 140)
            elif len(tk) >= 3 and _is_int(tk[0]) and _is_int(tk[1]):
                t1, a1, e2, wild = int(tk[0]), int(tk[1]), tk[2][:2], 1
            elif len(tk) >= 2:
                e1, e2, wild = tk[0][:2], tk[1][:2], 2
        else:
            if len(tk) >= 3 and all(_is_int(x) for x in tk[:3]):
                t1, a1, a2 = (int(x) for x in tk[:3]); t2 = t1; wild = 0
            elif len(tk) >= 3 and _is_int(tk[0]) and _is_int(tk[2]):
                t1, e2, a1, wild = int(tk[0]), tk[1][:2], int(tk[2]), 1          # swapped
            elif len(tk) >= 3 and _is_int(tk[0]) and _is_int(tk[1]):
                t1, a1, e2, wild = int(tk[0]), int(tk[1]), tk[2][:2], 1
            elif len(tk) >= 2:
                e1, e2, wild = tk[0][:2], tk[1][:2], 2
        if wild is None:
            break
        if wild == 0:
            ref, obs = [(t1, a1)], [(t2, a2)]
            desc = f"Atom #{a1} in molecule type {t1} -> atom #{a2} in molecule type {t2}"
            weight = type_nmol[t1 - 1]
        elif wild == 1:
            ref, obs = [(t1, a1)], atoms_of(e2, 0 if inter else t1)
            if not obs:
                continue
            desc = f"Atom #{a1} in molecule type {t1} -> {len(obs)} {e2} atoms."
            weight = type_nmol[t1 - 1]
        else:
            ref, obs = atoms_of(e1), atoms_of(e2)
            if not ref or not obs:
                continue
            desc = f"{len(ref)} {e1} atoms -> {len(obs)} {e2} atoms."
            weight = 0
            for x in ref:                                                        # count_wildcard_weights
                for y in obs:
                    if not inter:
                        if x[0] == y[0] and x[1] != y[1]:
                            weight += type_nmol[x[0] - 1]; break
                    elif x[0] != y[0] or x[1] != y[1]:
                        weight += type_nmol[x[0] - 1]; break
                    elif type_nmol[x[0] - 1] > 1:
                        weight += type_nmol[x[0] - 1]; break
        subs.append(dict(ref=ref, obs=obs, weight=weight, description=desc,
                         el1=type_elements[ref[0][0] - 1][ref[0][1] - 1], el2=type_elements[obs[0][0] - 1][obs[0][1] - 1]))
    kw = dict(nsteps=1, sampling_interval=1, calc_stdev=False, maxdist=10.0, exponent=1.0)
    calc_exp, calc_ffc = sticky.get("calc_exp", False), sticky.get("calc_ffc", False)
    for t in (_tokens(l) for l in lines[1 + nsub:]):
        if not t:
            continue
        k = t[0]
        try:
            if k == "sampling_interval":
                kw["sampling_interval"] = int(t[1])
            elif k == "exponent":
                kw["exponent"] = float(t[1])
            elif k in ("exponential_weight", "weigh_exponential", "calculate_exponential", "average_exponential"):
                calc_exp = _logical(t[1])
            elif k in ("stdev", "standard_deviation", "standarddev"):
                kw["calc_stdev"] = _logical(t[1])
            elif k in ("ffc_weight", "weigh_ffc", "calculate_ffc", "average_ffc", "ffc", "dhh", "rhh"):
                calc_ffc = _logical(t[1])
            elif k == "nsteps":
                kw["nsteps"] = int(t[1])
            elif k in ("maxdist", "cutoff"):
                kw["maxdist"] = float(t[1])
            elif k in ("maxdist_optimize", "cutoff_optimize"):
                import numpy as np
                kw["maxdist"] = float(np.float32(box_limit / 2.0)) if box_limit is not None else 10.0
            elif k == "quit":
                break
        except (IndexError, ValueError):
            pass
    sticky["calc_exp"], sticky["calc_ffc"] = calc_exp, calc_ffc
    return dict(inter=inter, subjobs=subs, calc_exp=calc_exp, calc_ffc=calc_ffc, **kw)


def make_dist_job(spec):
    """dict from parse_distance_input -> (pa_dist_job, keepalive)."""
    import ctypes as C
    import numpy as np
    n = len(spec["subjobs"])
    arr = (capi.pa_dist_subjob * n)()
    keep = []
    for i, sj in enumerate(spec["subjobs"]):
        ref = np.ascontiguousarray(np.array(sj["ref"], dtype=np.int32).reshape(-1, 2))
        obs = np.ascontiguousarray(np.array(sj["obs"], dtype=np.int32).reshape(-1, 2))
        keep += [ref, obs]
        arr[i].n_ref, arr[i].n_obs = len(ref), len(obs)
        arr[i].ref = ref.ctypes.data_as(C.POINTER(C.c_int32))
        arr[i].obs = obs.ctypes.data_as(C.POINTER(C.c_int32))
        arr[i].weight = int(sj["weight"])
    job = capi.pa_dist_job()
    job.intermolecular, job.nsubjobs = int(spec["inter"]), n
    job.subjobs = C.cast(arr, C.POINTER(capi.pa_dist_subjob))
    job.nsteps, job.sampling_interval = int(spec["nsteps"]), int(spec["sampling_interval"])
    job.calc_exp, job.calc_ffc, job.calc_stdev = int(spec["calc_exp"]), int(spec["calc_ffc"]), int(spec["calc_stdev"])
    import numpy as np
    job.maxdist, job.exponent = float(np.float32(spec["maxdist"])), float(np.float32(spec["exponent"]))
    keep.append(arr)
    return job, keep


def _e93(v):
    """E9.3: 0.dddE+ee with the correctly rounded digits of the binary value (exact ties go to the even digit)."""
    if v == 0.0:
        return "0.000E+00"
    mant, exp = ("%.2e" % abs(v)).split("e")
    return f"{'-' if v < 0 else ''}0.{mant.replace('.', '')}E{int(exp) + 1:+03d}"


def format_distance(d):
    """formatted_distance_output, Distances:1145-1157."""
    if d > 1000.0:
        return _e93(d)
    if d > 1.0:
        return f"{d:.3f}"
    if d > 0.01:
        return f"{d:5.3f}"
    return _e93(d)


def distance_report(spec, results):
    """The lines report_distances prints (Distances:1093-1143), for comparison with the reference's stdout."""
    out = []
    n = len(spec["subjobs"])
    l = "d" if spec["inter"] else "r"
    for i, (sj, r) in enumerate(zip(spec["subjobs"], results)):
        out.append(f"   Subjob #{i + 1} out of {n}:")
        out.append(f"     {sj['description']}")
        if r["missed_neighbours"] > 0:
            out.append(f"     {r['missed_neighbours']} atoms without neighbour - consider increasing cutoff.")
        else:
            out.append(f"     Used {sj['weight']} reference atoms per step.")
        out.append(f"     average closest distance: {format_distance(r['closest_average'])}")
        if spec["calc_stdev"]:
            out.append(f"     standard deviation: {format_distance(r['closest_stdev'])}")
        if spec["calc_exp"]:
            out.append(f"     exponentially weighed distance: {format_distance(r['exponential_average'])}")
            if spec["calc_stdev"]:
                out.append(f"     standard deviation: {format_distance(r['exponential_stdev'])}")
        if spec["calc_ffc"]:
            out.append(f"     FFC distance {l}{sj['el1']}{sj['el2']}: {format_distance(r['ffc_average'])}")
            out.append(f"     exp(-kr) weighed FFC distance {l}'{sj['el1']}{sj['el2']}: {format_distance(r['ffcexp_average'])}")
            out.append(f"     Number of averages for FFC: {format_distance(r['ffc_weight'])}")
    return out
