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
    if mode not in ("cdf", "pdf"):
        raise ValueError(99 if mode != "charge_arm" else -4)
    nref = int(head[1])
    refs = []
    for k in range(nref):
        t = _tokens(lines[1 + k])
        v = [int(x) for x in t[:5]]
        if len(v) < 5:
            raise ValueError(99)
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
                if mode == "pdf":
                    kw["subtract_uniform"] = _logical(t[1])
            elif key == "weigh_charge":
                kw["weigh_charge"] = _logical(t[1])
            elif key in ("use_coc", "use_COC", "center_of_charge", "centre_of_charge"):
                kw["use_coc"] = _logical(t[1])
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
                                      maxdist_optimize=kw["maxdist_optimize"], maxdist=kw["maxdist"]))
    return reqs
