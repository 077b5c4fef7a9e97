#!/usr/bin/env python
"""Differential check of the `distance` / `contact_distance` keywords of the drop-in against the reference binary on
random inputs (test infrastructure like make_golden.py: needs oracle/_ref, i.e. /root/reference, so it is run by hand
and not collected by pytest).  The drop-in runs unchanged in a subprocess with pa_distance_subjobs / pa_contact_distance
interposed by the oracle's loops (LD_PRELOAD of a small shim compiled here, no GPU needed); compared are the report
blocks both programs print (the analyses write no files).
usage: python tests/golden/fuzz_distance.py [ncases] [seed]"""
import os
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

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
orc.load()
assert orc.ref_available()
fx = fixtures.Fixture("re4")
natoms = {1: 5, 2: 6, 3: 1}
ELS = ["P", "O", "F", "C", "Li", "H"]

shim_dir = tempfile.mkdtemp(prefix="shim_")
open(os.path.join(shim_dir, "shim.c"), "w").write(r'''
#include <stdint.h>
#include <string.h>
#include <dlfcn.h>
#include "prealpha_b200.h"
static void *lib(void) { static void *h = 0; if (!h) h = dlopen("%s", RTLD_NOW); return h; }
int pa_distance_subjobs(const pa_traj *traj, const pa_dist_job *job, const pa_exec *exec, pa_dist_result *results)
{
    int (*f)(const pa_traj *, const pa_dist_job *, pa_dist_result *) = (int (*)(const pa_traj *, const pa_dist_job *, pa_dist_result *))dlsym(lib(), "orc_distance_subjobs");
    return f(traj, job, results);
}
int pa_contact_distance(const pa_traj *traj, int32_t timestep, int32_t type, const pa_exec *exec, pa_contact_result *result)
{
    int (*f)(const pa_traj *, int, int, int, float *) = (int (*)(const pa_traj *, int, int, int, float *))dlsym(lib(), "orc_contact_distance");
    float out[3];
    memset(result, 0, sizeof *result);
    int rc = f(traj, timestep, type, 4, out);
    result->largest_intramolecular = out[0]; result->smallest_intramolecular = out[1]; result->smallest_intermolecular = out[2];
    return rc;
}
''' % os.path.join(ROOT, "oracle", "libpa_oracle.so"))
subprocess.check_call(["gcc", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(shim_dir, "shim.so"),
                       os.path.join(shim_dir, "shim.c"), "-ldl"])


def random_input():
    inter = rng.random() < 0.5
    n = int(rng.integers(1, 5))
    text = f"{'intermolecular' if inter else 'intramolecular'} {n}\n"
    for _ in range(n):
        kind = rng.integers(0, 4)
        t1 = int(rng.choice([1, 2, 3])); t2 = int(rng.choice([1, 2, 3]))
        if kind == 0:
            if inter:
                text += f"{t1} {int(rng.integers(1, natoms[t1] + 1))} {t2} {int(rng.integers(1, natoms[t2] + 1))}\n"
            else:
                text += f"{t1} {int(rng.integers(1, natoms[t1] + 1))} {int(rng.integers(1, natoms[t1] + 1))}\n"
        elif kind == 1:
            text += f"{t1} {int(rng.integers(1, natoms[t1] + 1))} {rng.choice(ELS)}\n"
        elif kind == 2:
            text += f"{rng.choice(ELS)} {t1} {int(rng.integers(1, natoms[t1] + 1))}\n"
        else:
            text += f"{rng.choice(ELS)} {rng.choice(ELS)}\n"
    text += "cutoff_optimize\n" if rng.random() < 0.3 else f"maxdist {rng.uniform(2.0, 16.0):.2f}\n"
    text += f"nsteps {int(rng.integers(1, 13))}\nsampling_interval {int(rng.integers(1, 4))}\n"
    for kw in ("stdev", "ffc", "average_exponential"):
        if rng.random() < 0.5:
            text += f"{kw} {rng.choice(['T', 'F'])}\n"
    if rng.random() < 0.4:
        text += f"exponent {rng.uniform(0.5, 3.0):.2f}\n"
    return text + "quit\n"


def blocks(stdout):
    out, cur = [], None
    for line in stdout.splitlines():
        if "molecular distance calculation. Results:" in line or "contact distances" in line:
            if cur is not None and "contact distances" in line:
                out.append(cur)
            cur = [line.rstrip()] if "contact distances" in line else []
        elif line.startswith(" ( reference atom(s)") and cur is not None:
            out.append(cur); cur = None
        elif line.startswith(" elapsed time") or line.startswith(" Line "):
            if cur is not None and cur and "contact distances" in cur[0]:
                out.append(cur); cur = None
        elif cur is not None and "###" not in line:               # (the binary's timing lines of its parallel sections)
            cur.append(line.rstrip())
    return out


bad = skipped = 0
for case in range(ncases):
    texts = [random_input() for _ in range(2)]
    contact = f" contact_distance {int(rng.integers(1, 13))} {int(rng.choice([-1, 1, 2, 3]))}\n"
    dirs = []
    for who in ("ref", "mine"):
        wd = tempfile.mkdtemp(prefix=f"fuzzx_{who}_")
        os.makedirs(os.path.join(wd, "output"))
        orc.write_xyz(os.path.join(wd, "traj.xyz"), fx.elements, fx.coords.astype(np.float64), fmt="%.9g")
        open(os.path.join(wd, "molecular.inp"), "w").write(f" {fx.nsteps}\n {len(fx.types)}\n" + "".join(f" {c:+d} {a} {m}\n" for c, a, m in fx.types) + "quit\n")
        for k, t in enumerate(texts):
            open(os.path.join(wd, f"d{k}.inp"), "w").write(t)
        open(os.path.join(wd, "general.inp"), "w").write(' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./output/"\n set_threads 4\n cubic_box_edge 0.0 34.7853\n parallel_operation T\n'
                                                          ' sequential_read F\n trajectory_type xyz\n distance "d0.inp"\n' + contact + ' distance "d1.inp"\n quit\n')
        dirs.append(wd)
    try:
        out, _ = orc.run_reference(dirs[0])
    except Exception as e:
        skipped += 1
        print(f"case {case}: reference failed ({str(e)[:80]!r}), skipped")
        for d in dirs:
            shutil.rmtree(d)
        continue
    if "ERROR 0." in out:
        # the binary's own internal error: an `El1 El2` line whose first element has two letters is written to its scratch
        # file as "LiF" (list-directed output does not separate adjacent character items, Distances:650) and read back wrongly
        skipped += 1
        print(f"case {case}: the reference reports its internal error 0, skipped")
        for d in dirs:
            shutil.rmtree(d)
        continue
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from prealpha_b200 import capi, api; "
            "rc = capi.load().pa_run_general_input(b'general.inp', C.byref(api.make_exec())); print('RC', rc)" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], cwd=dirs[1], capture_output=True, env=dict(os.environ, LD_PRELOAD=os.path.join(shim_dir, "shim.so")))
    mine = p.stdout.decode()
    b_ref, b_mine = blocks(out), blocks(mine)
    if b_ref != b_mine:
        bad += 1
        keep = f"/tmp/fuzzx_fail_{case}"
        shutil.rmtree(keep, ignore_errors=True); os.makedirs(keep)
        open(keep + "/ref_stdout.txt", "w").write(out); open(keep + "/my_stdout.txt", "w").write(mine + p.stderr.decode())
        for k, t in enumerate(texts):
            open(keep + f"/d{k}.inp", "w").write(t)
        print(f"case {case}: MISMATCH (kept in {keep}) contact={contact.strip()}")
        for x, y in zip(b_ref, b_mine):
            if x != y:
                print("   ref :", x[:6]); print("   mine:", y[:6]); break
    else:
        print(f"case {case}: ok ({len(b_ref)} report blocks, {sum(len(b) for b in b_ref)} lines)")
    for d in dirs:
        shutil.rmtree(d)
shutil.rmtree(shim_dir)
print("mismatches:", bad, "of", ncases, "skipped:", skipped)
