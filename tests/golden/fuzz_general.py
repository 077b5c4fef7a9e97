#!/usr/bin/env python
"""Differential check of whole `general.inp` runs of the drop-in against the reference binary (test infrastructure like
make_golden.py: needs oracle/_ref, i.e. /root/reference, so it is run by hand and not collected by pytest): a random
sequence of the keywords of this path -- distribution / distribution_simple / charge_arm_simple / clm_simple / distance /
distance_simple / contact_distance(_simple) / cluster / speciation with random prefixes in between -- in ONE run, so that
what carries over from one keyword to the next (prefixes, the sticky switches of the distance module, the generated
prealpha_simple.inp, error counters) is covered.  All GPU entry points are interposed by the oracle's loops (LD_PRELOAD of
a shim compiled here); compared are all output files byte for byte and the final error / warning count.
usage: python tests/golden/fuzz_general.py [ncases] [seed]"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_lib as orc  # noqa: E402
import fixtures  # noqa: E402

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
orc.load()
assert orc.ref_available()
fx = fixtures.Fixture("re4")
natoms = {1: 5, 2: 6, 3: 1}

shim_dir = tempfile.mkdtemp(prefix="shim_")
open(os.path.join(shim_dir, "shim.c"), "w").write(r'''
#include <stdint.h>
#include <string.h>
#include <dlfcn.h>
#include "prealpha_b200.h"
static void *sym(const char *n) { static void *h = 0; if (!h) h = dlopen("%s", RTLD_NOW); return dlsym(h, n); }
int pa_distribution_histogram_multi(const pa_traj *traj, const pa_job *jobs, int32_t njobs, const pa_exec *exec,
                                    int64_t *const *hists, int64_t *within_bin, int64_t *out_of_bounds, pa_stats *stats)
{
    int (*f)(const pa_traj *, const pa_job *, int, int64_t *, int64_t *, int64_t *) = sym("orc_distribution_histogram");
    if (stats) memset(stats, 0, sizeof *stats);
    for (int j = 0; j < njobs; ++j) { int rc = f(traj, &jobs[j], 8, hists[j], &within_bin[j], &out_of_bounds[j]); if (rc) return rc; }
    return 0;
}
int pa_distance_subjobs(const pa_traj *traj, const pa_dist_job *job, const pa_exec *exec, pa_dist_result *results)
{
    int (*f)(const pa_traj *, const pa_dist_job *, pa_dist_result *) = sym("orc_distance_subjobs");
    return f(traj, job, results);
}
int pa_contact_distance(const pa_traj *traj, int32_t timestep, int32_t type, const pa_exec *exec, pa_contact_result *result)
{
    int (*f)(const pa_traj *, int, int, int, float *) = sym("orc_contact_distance");
    float out[3];
    memset(result, 0, sizeof *result);
    int rc = f(traj, timestep, type, 4, out);
    result->largest_intramolecular = out[0]; result->smallest_intramolecular = out[1]; result->smallest_intermolecular = out[2];
    return rc;
}
int pa_neighbour_pairs(const pa_traj *traj, const pa_neigh_request *rq, const pa_exec *exec, pa_neigh_pair *out, int64_t cap, int64_t *count, int64_t *tests)
{
    int (*f)(const pa_traj *, const pa_neigh_request *, pa_neigh_pair *, int64_t, int64_t *) = sym("orc_neighbour_pairs");
    if (tests) *tests = 1;
    return f(traj, rq, out, cap, count);
}
''' % os.path.join(ROOT, "oracle", "libpa_oracle.so"))
subprocess.check_call(["gcc", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(shim_dir, "shim.so"),
                       os.path.join(shim_dir, "shim.c"), "-ldl"])

VECS = ["+0 +0 +1", "+1 +0 +0", "+1 +1 +1", "+0 +1 -1"]


def dist_input():
    mode = str(rng.choice(["pdf", "pdf", "cdf", "charge_arm"]))
    n = int(rng.integers(1, 3))
    text = f"{mode} {n}\n"
    for _ in range(n):
        ref = int(rng.integers(1, 4)) if rng.random() < 0.93 else int(rng.choice([0, 4, 9]))      # (now and then an invalid molecule type)
        text += f"{rng.choice(VECS)} {ref}\n" if mode == "charge_arm" else f"{rng.choice(VECS)} {ref} {int(rng.choice([-1, 1, 2, 3]))}\n"
    text += "maxdist_optimize\n" if rng.random() < 0.5 else f"maxdist {rng.uniform(2.0, 20.0):.2f}\n"
    if rng.random() < 0.5:
        text += f"weigh_charge {rng.choice(['T', 'F'])}\n"
    return text + f"sampling_interval {int(rng.integers(1, 4))}\nbin_count_a {int(rng.integers(4, 60))}\nbin_count_b {int(rng.choice([1, int(rng.integers(2, 20))]))}\nquit\n"


def distance_input():
    inter = rng.random() < 0.5
    text = f"{'intermolecular' if inter else 'intramolecular'} 2\n"
    for _ in range(2):
        t1, t2 = int(rng.choice([1, 2])), int(rng.choice([1, 2, 3]))
        if rng.random() < 0.08:
            text += f"{t1} {int(rng.choice([0, 9]))} {t2} 1\n" if inter else f"{t1} 9 1\n"          # invalid atom index
        elif rng.random() < 0.5:
            text += (f"{t1} {int(rng.integers(1, natoms[t1] + 1))} {t2} {int(rng.integers(1, natoms[t2] + 1))}\n" if inter
                     else f"{t1} {int(rng.integers(1, natoms[t1] + 1))} {int(rng.integers(1, natoms[t1] + 1))}\n")
        else:
            text += f"{t1} {int(rng.integers(1, natoms[t1] + 1))} {rng.choice(['O', 'F', 'C', 'P'])}\n"
    text += f"maxdist {rng.uniform(3.0, 12.0):.2f}\nnsteps {int(rng.integers(1, 13))}\n"
    for kw in ("stdev", "ffc", "average_exponential"):
        if rng.random() < 0.5:
            text += f"{kw} {rng.choice(['T', 'F'])}\n"
    return text + "quit\n"


def cluster_input():
    if rng.random() < 0.5:
        text = "GLOBAL\n" + "".join(f"add_atom_index {t} {int(rng.integers(1, natoms[t] + 1))}\n" for t in rng.choice([1, 2, 3], size=2))
        text += f"single_cutoff {rng.uniform(2.0, 5.0):.2f}\n"
    else:
        text = "PAIRS\n" + "".join(f"pair {a} {int(rng.integers(1, natoms[a] + 1))} {b} {int(rng.integers(1, natoms[b] + 1))} {rng.uniform(2.0, 4.5):.2f}\n"
                                   for a, b in rng.choice([1, 2, 3], size=(2, 2)))
    return text + f"nsteps {int(rng.choice([int(rng.integers(2, 13)), 40]))}\nsampling_interval {int(rng.integers(1, 3))}\nprint_statistics T\nautocorrelation F\nquit\n"


def speciation_input():
    text = " 1 acceptor\n 3 1 0.0\n 2 donors\n" + f" 1 1 {rng.uniform(3.0, 4.5):.2f}\n 2 {int(rng.choice([2, 5]))} {rng.uniform(2.2, 3.2):.2f}\n"
    return text + f" N_species {int(rng.integers(5, 40))}\n N_neighbours {int(rng.integers(3, 12))}\n nsteps {int(rng.integers(2, 13))}\n" + (" use_ccc T\n" if rng.random() < 0.5 else "") + " quit\n"


def report_blocks(stdout):
    """distance / contact-distance report lines (these analyses write no files)."""
    keep = []
    for ln in stdout.splitlines():
        t = ln.strip()
        if "###" in ln:
            continue
        if (t.startswith(("Subjob #", "Used ", "average closest distance", "standard deviation:", "exponentially weighed distance", "FFC distance",
                          "Number of averages", "Largest intramolecular", "Smallest intramolecular", "Smallest intermolecular", "Molecule type index"))
                or "atoms ->" in t or "-> atom #" in t or "without neighbour" in t):
            keep.append(t)
    return keep


bad = skipped = 0
for case in range(ncases):
    files, body = {}, ""
    for k in range(int(rng.integers(3, 8))):
        kind = str(rng.choice(["distribution", "distribution_simple", "charge_arm_simple", "clm_simple", "distance", "distance_simple", "contact_distance",
                               "contact_distance_simple", "cluster", "speciation", "set_prefix", "set_prefix"]))
        if kind == "set_prefix":
            body += f' set_prefix "p{k}_"\n'
        elif kind in ("distribution_simple", "charge_arm_simple", "clm_simple", "distance_simple", "contact_distance_simple"):
            body += f" {kind}\n"
        elif kind == "contact_distance":
            body += f" contact_distance {int(rng.integers(1, 13))} {int(rng.choice([-1, 1, 2, 3]))}\n"
        else:
            name = f"in{k}.inp"
            files[name] = {"distribution": dist_input, "distance": distance_input, "cluster": cluster_input, "speciation": speciation_input}[kind]()
            body += f' {kind} "{name}"\n'
    dirs = []
    for who in ("ref", "mine"):
        wd = tempfile.mkdtemp(prefix=f"fuzzg_{who}_")
        os.makedirs(os.path.join(wd, "output"))
        orc.write_xyz(os.path.join(wd, "traj.xyz"), fx.elements, fx.coords.astype(np.float64), fmt="%.9g")
        open(os.path.join(wd, "molecular.inp"), "w").write(f" {fx.nsteps}\n {len(fx.types)}\n" + "".join(f" {c:+d} {a} {m}\n" for c, a, m in fx.types) + "quit\n")
        for n, t in files.items():
            open(os.path.join(wd, n), "w").write(t)
        open(os.path.join(wd, "general.inp"), "w").write(' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./output/"\n set_threads 8\n cubic_box_edge 0.0 34.7853\n parallel_operation F\n'
                                                          ' sequential_read F\n trajectory_type xyz\n' + body + " quit\n")
        dirs.append(wd)
    try:
        out, _ = orc.run_reference(dirs[0])
    except Exception as e:
        skipped += 1
        print(f"case {case}: reference failed ({str(e)[:60]!r}), skipped")
        for d in dirs:
            shutil.rmtree(d)
        continue
    if "ERROR 0." in out or "Encountered" not in out:
        # its own internal error, or a severe error ended the binary early (e.g. error 22 "couldn't allocate memory" in the
        # distribution analysis that follows one rejected with error 33: the arrays of the rejected one are still allocated)
        skipped += 1
        print(f"case {case}: the reference reports its internal error 0 or stopped early, skipped")
        for d in dirs:
            shutil.rmtree(d)
        continue
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from prealpha_b200 import capi, api; "
            "rc = capi.load().pa_run_general_input(b'general.inp', C.byref(api.make_exec())); print('RC', rc)" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], cwd=dirs[1], capture_output=True, env=dict(os.environ, LD_PRELOAD=os.path.join(shim_dir, "shim.so")))
    mine = p.stdout.decode()
    skip = ("_first.xyz", "_last.xyz", "autocorrelation", "time_series")          # outputs this path does not write
    ref_files = sorted(f for f in os.listdir(os.path.join(dirs[0], "output")) if not f.endswith(skip) and "autocorrelation" not in f)
    my_files = sorted(os.listdir(os.path.join(dirs[1], "output")))
    same_files = ref_files == my_files and all(open(os.path.join(dirs[0], "output", f), "rb").read() == open(os.path.join(dirs[1], "output", f), "rb").read() for f in ref_files)
    simple_ref = open(os.path.join(dirs[0], "prealpha_simple.inp")).read() if os.path.exists(os.path.join(dirs[0], "prealpha_simple.inp")) else None
    simple_mine = open(os.path.join(dirs[1], "prealpha_simple.inp")).read() if os.path.exists(os.path.join(dirs[1], "prealpha_simple.inp")) else None
    m_ref = re.search(r"Encountered (\d+) errors/warnings globally", out)
    m_mine = re.search(r"Encountered (\d+) errors/warnings globally", mine)
    same_count = bool(m_ref and m_mine and m_ref.group(1) == m_mine.group(1))
    same_report = report_blocks(out) == report_blocks(mine)
    if not (same_files and same_report and simple_ref == simple_mine and same_count):
        bad += 1
        keep = f"/tmp/fuzzg_fail_{case}"
        shutil.rmtree(keep, ignore_errors=True); os.makedirs(keep)
        shutil.copytree(dirs[0], keep + "/ref"); shutil.copytree(dirs[1], keep + "/mine")
        open(keep + "/ref_stdout.txt", "w").write(out); open(keep + "/my_stdout.txt", "w").write(mine + p.stderr.decode())
        diff = [f for f in ref_files if f in my_files and open(os.path.join(dirs[0], "output", f), "rb").read() != open(os.path.join(dirs[1], "output", f), "rb").read()]
        print(f"case {case}: MISMATCH (kept in {keep}) files={same_files} report={same_report} simple={simple_ref == simple_mine} "
              f"count={m_ref.group(1) if m_ref else None}/{m_mine.group(1) if m_mine else None} only_ref={sorted(set(ref_files) - set(my_files))} "
              f"only_mine={sorted(set(my_files) - set(ref_files))} differ={diff}\n{body}")
    else:
        print(f"case {case}: ok ({len(ref_files)} files, {len(report_blocks(out))} report lines, {m_ref.group(1)} warnings)")
    for d in dirs:
        shutil.rmtree(d)
shutil.rmtree(shim_dir)
print("mismatches:", bad, "of", ncases, "skipped:", skipped)
