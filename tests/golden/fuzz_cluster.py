#!/usr/bin/env python
"""Differential check of the `cluster` replay against the reference binary on random inputs (test infrastructure like
make_golden.py: needs oracle/_ref, i.e. /root/reference, so it is run by hand and not collected by pytest).

The drop-in's host logic runs unchanged in a subprocess; only its GPU primitive pa_neighbour_pairs is interposed
(LD_PRELOAD of a ten-line shim compiled here) by the oracle's literal neighbour loops, so that the comparison needs no GPU.
The `*cluster_statistics.dat` files of both programs are compared byte for byte.
usage: python tests/golden/fuzz_cluster.py [ncases] [seed]"""
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

shim_dir = tempfile.mkdtemp(prefix="shim_")
open(os.path.join(shim_dir, "shim.c"), "w").write(r'''
#include <stdint.h>
#include <dlfcn.h>
typedef int (*orc_fn)(const void *, const void *, void *, int64_t, int64_t *);
int pa_neighbour_pairs(const void *traj, const void *rq, const void *exec, void *out, int64_t cap, int64_t *count, int64_t *tests)
{
    static orc_fn f = 0;
    if (!f) { void *h = dlopen("%s", RTLD_NOW); f = (orc_fn)dlsym(h, "orc_neighbour_pairs"); }
    if (tests) *tests = 1;
    return f(traj, rq, out, cap, count);
}
''' % os.path.join(ROOT, "oracle", "libpa_oracle.so"))
subprocess.check_call(["gcc", "-shared", "-fPIC", "-o", os.path.join(shim_dir, "shim.so"), os.path.join(shim_dir, "shim.c"), "-ldl"])


def random_input():
    if rng.random() < 0.5:
        text = "GLOBAL\n"
        for _ in range(int(rng.integers(1, 5))):
            t = int(rng.choice([1, 2, 3]))
            if rng.random() < 0.25:
                text += f"add_molecule_type {t}\n"
            else:
                text += f"add_atom_index {t} {int(rng.integers(1, natoms[t] + 1))}\n"
        if rng.random() < 0.5:
            for _ in range(int(rng.integers(1, 4))):
                text += f"single_cutoff {rng.uniform(1.5, 6.0):.3f}\n"
        else:
            lo = rng.uniform(1.5, 3.5)
            text += f"scan_cutoff {lo:.2f} {lo + rng.uniform(0.5, 3.0):.2f} {int(rng.integers(2, 6))}\n"
    else:
        text = "PAIRS\n"
        for _ in range(int(rng.integers(1, 6))):
            t1, t2 = int(rng.choice([1, 2, 3])), int(rng.choice([1, 2, 3]))
            text += f"pair {t1} {int(rng.integers(1, natoms[t1] + 1))} {t2} {int(rng.integers(1, natoms[t2] + 1))} {rng.uniform(1.5, 5.0):.2f}\n"
    text += f"nsteps {int(rng.integers(1, 13))}\nsampling_interval {int(rng.integers(1, 4))}\nprint_statistics T\nautocorrelation F\nquit\n"
    return text


bad = 0
for case in range(ncases):
    text = random_input()
    dirs = []
    for who in ("ref", "mine"):
        wd = tempfile.mkdtemp(prefix=f"fuzzc_{who}_")
        os.makedirs(os.path.join(wd, "output"))
        orc.write_xyz(os.path.join(wd, "traj.xyz"), fx.elements, fx.coords.astype(np.float64), fmt="%.9g")
        open(os.path.join(wd, "molecular.inp"), "w").write(f" {fx.nsteps}\n {len(fx.types)}\n" + "".join(f" {c:+d} {a} {m}\n" for c, a, m in fx.types) + "quit\n")
        open(os.path.join(wd, "cl.inp"), "w").write(text)
        open(os.path.join(wd, "general.inp"), "w").write(' "traj.xyz"\n "molecular.inp"\n "./"\n "./"\n "./output/"\n cubic_box_edge 0.0 34.7853\n parallel_operation F\n'
                                                          ' sequential_read F\n trajectory_type xyz\n set_prefix "f_"\n cluster cl.inp\n quit\n')
        dirs.append(wd)
    try:
        out, _ = orc.run_reference(dirs[0])
    except Exception as e:                                     # the binary itself can fail on odd inputs: not a comparison
        print(f"case {case}: reference failed ({str(e)[:80]!r}), skipped")
        for d in dirs:
            shutil.rmtree(d)
        continue
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from prealpha_b200 import capi, api; "
            "rc = capi.load().pa_run_general_input(b'general.inp', C.byref(api.make_exec())); print('RC', rc)" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], cwd=dirs[1], capture_output=True, env=dict(os.environ, LD_PRELOAD=os.path.join(shim_dir, "shim.so")))
    ref_files = sorted(f for f in os.listdir(os.path.join(dirs[0], "output")) if "statistics" in f)
    my_files = sorted(f for f in os.listdir(os.path.join(dirs[1], "output")) if "statistics" in f)
    ok = ref_files == my_files and all(open(os.path.join(dirs[0], "output", f), "rb").read() == open(os.path.join(dirs[1], "output", f), "rb").read() for f in ref_files)
    if not ok:
        bad += 1
        keep = f"/tmp/fuzzc_fail_{case}"
        shutil.rmtree(keep, ignore_errors=True); os.makedirs(keep)
        shutil.copytree(dirs[0], keep + "/ref"); shutil.copytree(dirs[1], keep + "/mine")
        open(keep + "/ref_stdout.txt", "w").write(out); open(keep + "/my_stdout.txt", "w").write(p.stdout.decode() + p.stderr.decode())
        print(f"case {case}: MISMATCH (kept in {keep}); files ref={ref_files} mine={my_files}\n{text}")
    else:
        print(f"case {case}: ok ({len(ref_files)} files)")
    for d in dirs:
        shutil.rmtree(d)
shutil.rmtree(shim_dir)
print("mismatches:", bad, "of", ncases)
