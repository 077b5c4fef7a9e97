#!/usr/bin/env python
"""Differential check of the drop-in's STREAMING distribution path (`sequential_read T`: batches of frames read from the
file by the reader thread, dealt to a session by the worker thread, sampled frames picked per batch, short files) against
the reference binary on random inputs (test infrastructure like make_golden.py: needs oracle/_ref, i.e.
/root/reference, so it is run by hand and not collected by pytest).

The drop-in runs unchanged in a subprocess; only the session entry points are interposed (LD_PRELOAD of a small shim
compiled here: a session that keeps the uploaded frames on the host and evaluates them with the oracle's loops), so that the comparison needs no GPU: what is compared are ALL
output files of both programs, byte for byte -- i.e. everything between general.inp and the files except the CUDA
kernels, which the -m gpu tests pin on the oracle's integers.
usage: python tests/golden/fuzz_streaming.py [ncases] [seed]"""
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
FIX = {"re4": fixtures.Fixture("re4"), "lmp": fixtures.Fixture("lmp")}

shim_dir = tempfile.mkdtemp(prefix="shim_")
open(os.path.join(shim_dir, "shim.c"), "w").write(r'''
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include "prealpha_b200.h"
/* a session that keeps the uploaded frames on the host and evaluates them with the oracle's loops */
struct pa_session { pa_traj meta; pa_type *types; pa_job *jobs; int njobs, max_frames, nframes; float **buf; int64_t **hist; int64_t *within, *oob; size_t *nb; int64_t frames_done; };
typedef int (*orc_fn)(const pa_traj *, const pa_job *, int, int64_t *, int64_t *, int64_t *);
static orc_fn orc(void) { static orc_fn f = 0; if (!f) { void *h = dlopen("%s", RTLD_NOW); f = (orc_fn)dlsym(h, "orc_distribution_histogram"); } return f; }
int pa_device_count(void) { return 1; }
int pa_session_create(const pa_traj *m, const pa_job *jobs, int32_t njobs, const pa_exec *exec, int32_t max_frames, pa_session **out)
{
    pa_session *s = calloc(1, sizeof *s);
    s->meta = *m; s->types = malloc(sizeof(pa_type) * m->ntypes); memcpy(s->types, m->types, sizeof(pa_type) * m->ntypes); s->meta.types = s->types;
    s->jobs = malloc(sizeof(pa_job) * njobs); memcpy(s->jobs, jobs, sizeof(pa_job) * njobs); s->njobs = njobs; s->max_frames = max_frames;
    s->buf = calloc(m->ntypes, sizeof(float *));
    for (int t = 0; t < m->ntypes; ++t) s->buf[t] = malloc(sizeof(float) * 3 * (size_t)m->types[t].natoms * m->types[t].nmol * max_frames);
    s->hist = calloc(njobs, sizeof(int64_t *)); s->within = calloc(njobs, 8); s->oob = calloc(njobs, 8); s->nb = calloc(njobs, sizeof(size_t));
    for (int j = 0; j < njobs; ++j) { s->nb[j] = (size_t)jobs[j].bin_a * jobs[j].bin_b; s->hist[j] = calloc(s->nb[j], 8); }
    *out = s; return 0;
}
int pa_session_upload(pa_session *s, const pa_traj *traj, int32_t first, int32_t nframes, int32_t stride)
{
    if (nframes < 1 || nframes > s->max_frames) return PA_ERR_BAD_ARG;
    for (int t = 0; t < s->meta.ntypes; ++t) {
        const size_t ff = 3 * (size_t)s->types[t].natoms * s->types[t].nmol;
        for (int f = 0; f < nframes; ++f) memcpy(s->buf[t] + (size_t)f * ff, traj->types[t].xyz + (size_t)(first + (size_t)f * stride) * ff, ff * sizeof(float));
    }
    s->nframes = nframes; return 0;
}
int pa_session_run(pa_session *s)
{
    pa_traj t = s->meta; pa_type *ty = malloc(sizeof(pa_type) * t.ntypes); memcpy(ty, s->types, sizeof(pa_type) * t.ntypes);
    for (int k = 0; k < t.ntypes; ++k) ty[k].xyz = s->buf[k];
    t.types = ty; t.nsteps = s->nframes;
    for (int j = 0; j < s->njobs; ++j) {
        pa_job jb = s->jobs[j]; jb.sampling_interval = 1;             /* the uploaded frames ARE the sampled ones */
        int64_t *h = calloc(s->nb[j], 8), w = 0, o = 0;
        int rc = orc()(&t, &jb, 8, h, &w, &o);
        if (rc) return rc;
        for (size_t b = 0; b < s->nb[j]; ++b) s->hist[j][b] += h[b];
        s->within[j] += w; s->oob[j] += o; free(h);
    }
    s->frames_done += s->nframes; free(ty); return 0;
}
int pa_session_download(pa_session *s, int64_t *const *hists, int64_t *within, int64_t *oob, int32_t reset)
{
    for (int j = 0; j < s->njobs; ++j) {
        if (hists && hists[j]) memcpy(hists[j], s->hist[j], s->nb[j] * 8);
        if (within) within[j] = s->within[j];
        if (oob) oob[j] = s->oob[j];
    }
    return 0;
}
/* (wrap_trajectory needs the whole trajectory in RAM: the drop-in then takes the one-shot entry point) */
int pa_distribution_histogram_multi(const pa_traj *traj, const pa_job *jobs, int32_t njobs, const pa_exec *exec,
                                    int64_t *const *hists, int64_t *within_bin, int64_t *out_of_bounds, pa_stats *stats)
{
    if (stats) memset(stats, 0, sizeof *stats);
    for (int j = 0; j < njobs; ++j) { int rc = orc()(traj, &jobs[j], 8, hists[j], &within_bin[j], &out_of_bounds[j]); if (rc) return rc; }
    return 0;
}
int pa_session_stats(pa_session *s, pa_stats *st) { memset(st, 0, sizeof *st); st->frames_done = s->frames_done; return 0; }
void pa_session_destroy(pa_session *s) { (void)s; }
''' % os.path.join(ROOT, "oracle", "libpa_oracle.so"))
subprocess.check_call(["gcc", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(shim_dir, "shim.so"),
                       os.path.join(shim_dir, "shim.c"), "-ldl"])

VECS = ["+0 +0 +1", "+1 +0 +0", "+1 +1 +1", "+0 +1 -1", "-2 +1 +3", "+0 +0 -1"]


def random_input(fx):
    mode = str(rng.choice(["pdf", "pdf", "cdf", "charge_arm"]))
    n = int(rng.integers(1, 4))
    text = f"{mode} {n}\n"
    nt = len(fx.types)
    for _ in range(n):
        ref = int(rng.integers(1, nt + 1))
        if mode == "charge_arm":
            text += f"{rng.choice(VECS)} {ref}\n"
        else:
            text += f"{rng.choice(VECS)} {ref} {int(rng.choice([-1] + list(range(1, nt + 1))))}\n"
    if rng.random() < 0.5:
        text += "maxdist_optimize\n"
    else:
        text += f"maxdist {rng.uniform(1.0, 25.0):.2f}\n"
    if rng.random() < 0.5:
        text += f"subtract_uniform {rng.choice(['T', 'F'])}\n"
    if rng.random() < 0.5:
        text += f"weigh_charge {rng.choice(['T', 'F'])}\n"
    if rng.random() < 0.3:
        text += "use_coc F\n"        # (T needs atomic charges in every molecule: a neutral one makes the binary abort with error 94)
    if mode == "charge_arm" and rng.random() < 0.5:
        text += f"normalise_CLM {rng.choice(['T', 'F'])}\n"
    text += f"sampling_interval {int(rng.integers(1, 4))}\nbin_count_a {int(rng.integers(3, 90))}\nbin_count_b {int(rng.choice([1, 1, int(rng.integers(2, 30))]))}\nquit\n"
    return text


bad = skipped = 0
for case in range(ncases):
    fxname = str(rng.choice(["re4", "lmp"]))
    fx = FIX[fxname]
    text = random_input(fx)
    wrap = " wrap_trajectory T\n" if rng.random() < 0.3 else ""
    # (a file with fewer frames than molecular.inp declares is not compared: with sequential_read T the binary only notices at
    #  the end of the analysis -- severe error 85, or integrals normalised with the declared count -- while the drop-in
    #  reports error 84 and uses the frames that exist, as the binary itself does with sequential_read F, Molecular:5003-5009)
    short = False
    # optional sections of molecular.inp: element masses / charges, masses / charges of single atoms (centres of mass, charge arms)
    tail = ""
    els = sorted(set(fx.elements))
    if rng.random() < 0.4:
        pick = [e for e in els if rng.random() < 0.5] or els[:1]
        tail += f"masses {len(pick)}\n" + "".join(f" {e} {rng.uniform(1.0, 40.0):.3f}\n" for e in pick)
    if rng.random() < 0.4:
        rows = [(t + 1, int(rng.integers(1, fx.types[t][1] + 1))) for t in range(len(fx.types)) if rng.random() < 0.6]
        if rows:
            tail += f"atomic_masses {len(rows)}\n" + "".join(f" {t} {a} {rng.uniform(0.5, 50.0):.2f}\n" for t, a in rows)
    if rng.random() < 0.4:
        pick = [e for e in els if rng.random() < 0.5] or els[:1]
        tail += f"charges {len(pick)}\n" + "".join(f" {e} {rng.uniform(-1.5, 1.5):.3f}\n" for e in pick)
    if rng.random() < 0.5:
        rows = [(t + 1, a) for t in range(len(fx.types)) for a in range(1, fx.types[t][1] + 1) if rng.random() < 0.5]
        if rows:
            tail += f"atomic_charges {len(rows)}\n" + "".join(f" {t} {a} {rng.uniform(-1.2, 1.2):.4f}\n" for t, a in rows)
    dirs = []
    for who in ("ref", "mine"):
        wd = tempfile.mkdtemp(prefix=f"fuzzd_{who}_")
        os.makedirs(os.path.join(wd, "output"))
        if fxname == "re4":
            orc.write_xyz(os.path.join(wd, "traj.xyz"), fx.elements, fx.coords[:11 if short else None].astype(np.float64), fmt="%.9g")
            tname, box, ttype = "traj.xyz", " cubic_box_edge 0.0 34.7853\n", "xyz"
        else:
            orc.write_lmp(os.path.join(wd, "traj.lmp"), fx.elements, fx.coords.astype(np.float64), fx.box_lo, fx.box_hi, fmt="%.9g")
            tname, box, ttype = "traj.lmp", "", "lmp"
        open(os.path.join(wd, "molecular.inp"), "w").write(f" {fx.nsteps}\n {len(fx.types)}\n" + "".join(f" {c:+d} {a} {m}\n" for c, a, m in fx.types) + tail + "quit\n")
        open(os.path.join(wd, "dist.inp"), "w").write(text)
        open(os.path.join(wd, "general.inp"), "w").write(f' "{tname}"\n "molecular.inp"\n "./"\n "./"\n "./output/"\n set_threads 8\n{box} parallel_operation T\n'
                                                          f' sequential_read T\n{wrap} trajectory_type {ttype}\n set_prefix "f_"\n distribution dist.inp\n quit\n')
        dirs.append(wd)
    try:
        out, _ = orc.run_reference(dirs[0])
    except Exception as e:
        skipped += 1
        print(f"case {case}: reference failed ({str(e)[:80]!r}), skipped")
        for d in dirs:
            shutil.rmtree(d)
        continue
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from prealpha_b200 import capi, api; "
            "rc = capi.load().pa_run_general_input(b'general.inp', C.byref(api.make_exec())); print('RC', rc)" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], cwd=dirs[1], capture_output=True, env=dict(os.environ, LD_PRELOAD=os.path.join(shim_dir, "shim.so"), PA_STREAM_BATCH=str(int(rng.integers(1, 6)))))
    ref_files = sorted(os.listdir(os.path.join(dirs[0], "output")))
    my_files = sorted(os.listdir(os.path.join(dirs[1], "output")))
    ok = ref_files == my_files and all(open(os.path.join(dirs[0], "output", f), "rb").read() == open(os.path.join(dirs[1], "output", f), "rb").read() for f in ref_files)
    if not ok:
        bad += 1
        keep = f"/tmp/fuzzd_fail_{case}"
        shutil.rmtree(keep, ignore_errors=True); os.makedirs(keep)
        shutil.copytree(dirs[0], keep + "/ref"); shutil.copytree(dirs[1], keep + "/mine")
        open(keep + "/ref_stdout.txt", "w").write(out); open(keep + "/my_stdout.txt", "w").write(p.stdout.decode() + p.stderr.decode())
        diff = [f for f in ref_files if f in my_files and open(os.path.join(dirs[0], "output", f), "rb").read() != open(os.path.join(dirs[1], "output", f), "rb").read()]
        print(f"molecular.inp tail:\n{tail}")
        print(f"case {case}: MISMATCH [{fxname}{' wrap' if wrap else ''}{' short' if short else ''}] (kept in {keep}); only ref={sorted(set(ref_files) - set(my_files))} only mine={sorted(set(my_files) - set(ref_files))} differ={diff}\n{text}")
    else:
        errs = sorted({ln.strip()[:60] for ln in out.splitlines() if "ERROR" in ln})
        print(f"case {case}: ok ({len(ref_files)} files)" + (f" [{text.split()[0]}; reference said: {errs}]" if not ref_files else ""))
    for d in dirs:
        shutil.rmtree(d)
shutil.rmtree(shim_dir)
print("mismatches:", bad, "of", ncases, "skipped:", skipped)
