#!/usr/bin/env python
"""bench.py -- headline benchmark of the pair-distribution hot path.

Metric (BASELINE.json): binned atom-pairs/s on the synthetic 1M-atom cubic box (SURVEY 8d "C4"):
two one-atom types of 500 000 (Na, Cl), L = 271.4418 A (density 0.05 A^-3), xyz semantics,
references `+0 +0 +1 1 -1` and `+0 +0 +1 2 -1`, maxdist_optimize (cutoff L/2), 1000x1 bins.
One step = one frame per GPU = 1e12 ordered pair evaluations, ~5.2e11 of them binned.
Frames are sharded over the GPUs (weak scaling: one frame per GPU per step); the int64 bin
counts are combined once with an NCCL reduce at the end of the timed region.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun)
    python bench.py --impl reference ...                    (reference's own CPU build, bounded sample)

One JSON line on stdout (rank 0).  `value` is device-resident throughput, `e2e` goes through the
C ABI with host buffers (pinned staging + H2D + kernels + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_BOX = "271.4418"
N_PER_TYPE = 500_000
BINS = 1000
FP64_OPS_PER_PAIR = 26          # separable exact minimum image, non-FMA instruction issues
METRIC = "binned atom-pairs/s, 1M-atom RDF, 1/2/4/8 B200 vs OpenMP host cores"


def synth_frame(frame_index: int, natoms: int, L: float, seed: int = 12345) -> np.ndarray:
    """Counter-based generator keyed (seed, frame): u*L rounded to 5 decimals, then float32, so a text
    trajectory written with %.5f and a direct binary feed give identical bits."""
    g = np.random.Generator(np.random.Philox(key=[seed, frame_index]))
    return np.round(g.random((natoms, 3)) * L, 5).astype(np.float32)


def make_traj(frames: np.ndarray, n_per_type: int, L: float):
    from prealpha_b200 import capi
    nsteps = frames.shape[0]
    lo, hi = capi.cubic_box_edge(0.0, L)
    types = [dict(charge=+1, natoms=1, nmol=n_per_type, elements=["Na"], xyz=frames[:, :n_per_type].reshape(nsteps, n_per_type, 1, 3)),
             dict(charge=-1, natoms=1, nmol=n_per_type, elements=["Cl"], xyz=frames[:, n_per_type:].reshape(nsteps, n_per_type, 1, 3))]
    return capi.Trajectory(types, lo, hi, xyz_format=True)


def make_jobs(traj):
    from prealpha_b200 import api, capi
    return [api.job_setup(traj, capi.make_request(mode="pdf", vec=(0, 0, 1), ref_type=t, obs_type=-1, bin_a=BINS, bin_b=1,
                                                  maxdist_optimize=True, sampling_interval=1)) for t in (1, 2)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "500",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi's own start-up (driver attach) can stall CUDA calls for a few hundred ms: wait for its first
            # sample so that this happens before the timed region, not inside it
            t_end = time.time() + 8.0
            while not self.rows and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        # samples under load only (upper half of the power readings)
        if sm:
            thr = np.median(power) if len(power) > 2 else 0
            load = [c for c, p in zip(sm, power) if p >= thr] or sm
            return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                    "samples": len(sm), "reasons": sorted(reasons)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's OWN shipped binary on a bounded sample of the same
# generator (brute force: cost per pair is size independent), all host threads.

def cpu_reference_run(n_atoms_half: int, nframes: int, threads: int, workdir: str | None = None):
    """Returns dict(binned_per_s, pair_evals_per_s, seconds, kind, within)."""
    from oracle import oracle_lib as orc
    # same density as the 1M-atom box: scale the edge so that N/L^3 = 0.05
    n = 2 * n_atoms_half
    edge = round((n / 0.05) ** (1.0 / 3.0), 4)
    frames = np.stack([np.round(np.random.Generator(np.random.Philox(key=[12345, 10_000 + f])).random((n, 3)) * edge, 5)
                       for f in range(nframes)])
    pair_evals = nframes * 2 * (n_atoms_half * (n - 1))
    if orc.ref_probe():
        wd = workdir or tempfile.mkdtemp(prefix="pa_ref_")
        try:
            orc.write_xyz(os.path.join(wd, "traj.xyz"), ["Na"] * n_atoms_half + ["Cl"] * n_atoms_half, frames)
            text = f"pdf 2\n+0 +0 +1 1 -1\n+0 +0 +1 2 -1\nmaxdist_optimize\nsampling_interval 1\nbin_count_a {BINS}\nbin_count_b 1\nquit\n"
            orc.write_case(wd, "traj.xyz", [(+1, 1, n_atoms_half), (-1, 1, n_atoms_half)], nframes, ("0.0", f"{edge}"),
                           {"rdf.inp": text}, threads=threads)
            out, wall = orc.run_reference(wd)
            took = orc.parse_took_seconds(out)
            within = sum(w for w, _ in orc.parse_within(out))
            # the binary prints its own time with one decimal; prefer wall clock of the analysis when too coarse
            sec = sum(took) if took and sum(took) >= 2.0 else wall
            return dict(binned_per_s=within / sec, pair_evals_per_s=pair_evals / sec, seconds=sec, kind="reference",
                        within=within, threads=threads, atoms=n, frames=nframes)
        finally:
            if workdir is None:
                shutil.rmtree(wd, ignore_errors=True)
    # the shipped binary does not start on this box: time the C restatement instead (labelled "port")
    from prealpha_b200 import capi
    lo, hi = capi.cubic_box_edge(0.0, edge)
    fr32 = frames.astype(np.float32)
    types = [dict(charge=+1, natoms=1, nmol=n_atoms_half, elements=["Na"], xyz=fr32[:, :n_atoms_half].reshape(nframes, n_atoms_half, 1, 3)),
             dict(charge=-1, natoms=1, nmol=n_atoms_half, elements=["Cl"], xyz=fr32[:, n_atoms_half:].reshape(nframes, n_atoms_half, 1, 3))]
    traj = capi.Trajectory(types, lo, hi)
    t0 = time.time()
    within = 0
    for t in (1, 2):
        job = orc.job_setup(traj, capi.make_request(mode="pdf", ref_type=t, obs_type=-1, bin_a=BINS, bin_b=1, maxdist_optimize=True,
                                                    sampling_interval=1))
        _, w, _ = orc.histogram(traj, job, nthreads=threads)
        within += w
    sec = time.time() - t0
    return dict(binned_per_s=within / sec, pair_evals_per_s=pair_evals / sec, seconds=sec, kind="port", within=within,
                threads=threads, atoms=n, frames=nframes)


def cpu_sample_size(threads: int, target_seconds: float):
    """frames = 2 x threads (balanced STATIC,1 schedule, >= 11 needed by the module); atoms so that the run
    takes about target_seconds at ~2e6 pair-evals/s/thread of the -O0 reference build."""
    frames = max(12, 2 * threads)
    rounds = -(-frames // threads)
    pairs_per_frame = target_seconds * 2.0e6 / rounds
    n_half = int(round((pairs_per_frame / 4.0) ** 0.5))      # 2 refs x n_half x 2 n_half pairs
    return max(500, min(10000, n_half)), frames


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_half, frames = cpu_sample_size(threads, 6.0)
    for _ in range(args.warmup):
        cpu_reference_run(max(500, n_half // 2), frames, threads)
    res = []
    t0 = time.time()
    for _ in range(args.steps):
        res.append(cpu_reference_run(n_half, frames, threads))
    wall = time.time() - t0
    sec = sum(r["seconds"] for r in res)
    value = sum(r["within"] for r in res) / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "binned pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",     # the reference computes in fp64 (27 images per pair)
        "config": {"workload": f"synthetic cubic box (same generator and density as the 1M-atom case), {2 * n_half} atoms x {frames} frames, "
                               "references 1 -1 and 2 -1, maxdist_optimize, 1000x1 bins; brute force, cost per pair independent of N"},
        "pair_evals_per_s": sum(r["pair_evals_per_s"] * r["seconds"] for r in res) / sec,
        "cpu_baseline": {"value": value, "unit": "binned pairs/s", "cores": threads, "kind": res[0]["kind"],
                         "sample": f"{2 * n_half} atoms x {frames} frames per step, {args.steps} steps, "
                                   + ("shipped prealpha binary (gfortran -fopenmp, -O0), set_threads = all host cores" if res[0]["kind"] == "reference"
                                      else "C restatement oracle/pa_oracle.c (-O2, OpenMP), the shipped binary did not start here")},
        "e2e": {"value": value, "unit": "binned pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
def _traffic_from_profiles():
    """dram bytes (read + write) per launch of the dominant kernel, from the committed ncu --set full capture of THIS
    round's kernels (profiles/r2_traffic.json, written by tools/ncu_summary.py); None when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_launch"]), t.get("source")
    except Exception:
        return None, None


def make_c5(natoms: int, maxdist: float):
    """BASELINE config 5: one 4M-atom frame, 200x90 distance-angle histogram (`pdf` with 90 angle bins)."""
    from prealpha_b200 import api, capi
    L = round((natoms / 0.05) ** (1.0 / 3.0), 3)
    g = np.random.Generator(np.random.Philox(key=[12345, 5_000_000]))
    frame = np.round(g.random((natoms, 3)) * L, 5).astype(np.float32)
    half = natoms // 2
    lo, hi = capi.cubic_box_edge(0.0, L)
    types = [dict(charge=+1, natoms=1, nmol=half, elements=["Na"], xyz=frame[:half].reshape(1, half, 1, 3)),
             dict(charge=-1, natoms=1, nmol=half, elements=["Cl"], xyz=frame[half:].reshape(1, half, 1, 3))]
    traj = capi.Trajectory(types, lo, hi, xyz_format=True)
    jobs = [api.job_setup(traj, capi.make_request(mode="pdf", vec=(0, 0, 1), ref_type=1, obs_type=-1, bin_a=200, bin_b=90,
                                                  maxdist=maxdist, sampling_interval=1))]
    return traj, jobs, L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c4", choices=["c4", "c5"],
                    help="c4 (default, the BASELINE metric): 1M-atom RDF, one frame per GPU per step (weak scaling over frames); "
                         "c5: ONE 4M-atom frame, 200x90 distance-angle histogram, tile shards over the GPUs (strong scaling)")
    ap.add_argument("--atoms-per-type", type=int, default=N_PER_TYPE, help="debug: smaller box (not a valid bench line)")
    ap.add_argument("--c5-maxdist", type=float, default=40.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from prealpha_b200 import api, capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, (world, args.gpus)
    if not torch.cuda.is_available() or api.device_count() < 1:
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K = max(args.warmup, 3), args.steps
    c5 = args.config == "c5"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if c5:
        # strong scaling: every rank holds the SAME frame and takes tile shard `rank` of `world` (pa_exec.tile_shard_*)
        natoms = 4_000_000 if args.atoms_per_type == N_PER_TYPE else 2 * args.atoms_per_type
        traj1, jobs, L = make_c5(natoms, args.c5_maxdist)
        n_half = natoms // 2
        frames_of = lambda s: 0                                   # noqa: E731  (one resident frame, re-run every step)
        nres = 1
        traj = traj1
        ex = api.make_exec(device=local, flags=args.flags, tile_shard_rank=rank, tile_shard_count=world)
    else:
        n_half = args.atoms_per_type
        natoms = 2 * n_half
        L = float(L_BOX) if n_half == N_PER_TYPE else round((natoms / 0.05) ** (1.0 / 3.0), 4)
        # frames: step s of rank r is global frame s*world + r (frames sharded round-robin over the GPUs)
        frames = np.stack([synth_frame(s * world + rank, natoms, L) for s in range(W + K)])
        traj = make_traj(frames, n_half, L)
        jobs = make_jobs(traj)
        frames_of = lambda s: s                                   # noqa: E731
        nres = W + K
        ex = api.make_exec(device=local, flags=args.flags)

    # ---- device-resident arm (`value`): all frames uploaded before the timed region -------------------
    sess = api.Session(traj, jobs, max_frames=nres, exec_=ex)
    sess.upload(traj, 0, nres)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")    # 256 MB > 126 MB of L2
    for s in range(W):
        flush.zero_()                      # also warms torch's fill kernel: no cold launch inside the timed region
        torch.cuda.synchronize()
        sess.run(frames_of(s), 1)
    sess.download(reset=True)
    st0 = sess.stats()
    fp64_rate = api.fp64_issue_rate(local)
    atomic_rate = api.smem_atomic_rate(local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    flush.zero_()
    barrier()
    t0 = time.perf_counter()
    step_marks = []
    for s in range(K):
        flush.zero_()                      # L2 flush between timed iterations (0.04 ms of HBM writes on torch's stream)
        torch.cuda.synchronize()
        step_marks.append(time.perf_counter() - t0)
        sess.run(frames_of(W + s), 1)      # prepare + sort + pair kernels on the session's compute stream
    ptr, nwords = sess.device_accumulators()   # synchronises the compute stream
    step_marks.append(time.perf_counter() - t0)
    if world > 1:
        class _Acc:   # zero-copy view of the session's int64 accumulators for NCCL
            __cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        acc = torch.as_tensor(_Acc(), device=f"cuda:{local}")
        dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)      # integer bin counts combined once over NVLink
    barrier()
    t_value = time.perf_counter() - t0
    if rank == 0:
        print("[bench] value arm: step boundaries (s) " + " ".join(f"{m:.3f}" for m in step_marks) + f" end {t_value:.3f}", file=sys.stderr)
        gaps = np.diff(step_marks)
        if len(gaps) > 2 and gaps[0] > 1.5 * np.median(gaps[1:]):
            print(f"[bench] WARNING: first timed step took {gaps[0]:.3f} s against a median of {np.median(gaps[1:]):.3f} s", file=sys.stderr)
    hists, within, oob = sess.download()
    st1 = sess.stats()
    sess.close()
    d = lambda name: getattr(st1, name) - getattr(st0, name)      # noqa: E731
    kernel_ms, launches, pair_evals_rank = d("kernel_ms"), d("kernel_launches"), d("pairs_evaluated")
    tv = torch.tensor([t_value, kernel_ms, d("dominant_ms")], dtype=torch.float64, device="cuda")
    pe = torch.tensor([float(pair_evals_rank), float(d("cell_pair_evals")), float(d("smem_atomics")), float(d("unique_pairs_counted")),
                       float(d("dominant_pair_evals")), float(d("dominant_atomics")), float(d("dominant_launches")), float(launches)],
                      dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        dist.all_reduce(pe, op=dist.ReduceOp.SUM)
    t_value_max, kernel_ms_max, dominant_ms_max = (float(x) for x in tv)
    binned_total = int(within.sum())           # rank 0: whole job (after the reduce)
    pair_evals_total = float(pe[0])

    # ---- end-to-end arm: host buffers through the C ABI, H2D/D2H inside the timed region --------------
    if c5:
        e2e_traj, e2e_jobs, e2e_steps = traj, jobs, K
    else:
        e2e_traj = make_traj(frames[W:], n_half, L)
        e2e_jobs = make_jobs(e2e_traj)
        e2e_steps = 1
    mk = lambda: api.make_exec(device=local, frames_per_batch=1, flags=args.flags,                       # noqa: E731
                               tile_shard_rank=rank if c5 else 0, tile_shard_count=world if c5 else 1)
    warm_traj = traj if c5 else make_traj(frames[:1], n_half, L)
    api.histogram_multi(warm_traj, e2e_jobs, mk())        # warm: context, module load, page-locked staging (kept in the library's pool)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    w2_total = 0
    for _ in range(e2e_steps):
        h2, w2, o2, st_e = api.histogram_multi(e2e_traj, e2e_jobs, mk())
        h2d += st_e.h2d_bytes
        d2h += st_e.d2h_bytes
        if world > 1:
            acc2 = torch.from_numpy(np.concatenate([np.concatenate([h2[k], w2[k:k + 1], o2[k:k + 1]]) for k in range(len(e2e_jobs))])).cuda()
            dist.reduce(acc2, dst=0, op=dist.ReduceOp.SUM)
            nb = [int(j.bin_a) * int(j.bin_b) for j in e2e_jobs]
            flat = acc2.cpu().numpy()
            off = 0
            for k in range(len(e2e_jobs)):
                w2_total += int(flat[off + nb[k]])
                off += nb[k] + 2
        else:
            w2_total += int(w2.sum())
    barrier()
    t_e2e = time.perf_counter() - t0
    # the sampler runs through both timed regions: stopping nvidia-smi detaches it from the driver, which can stall
    # CUDA calls of this process for a few hundred ms -- not something to have between the two arms
    clocks = sampler.stop() if rank == 0 else None
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e_max = float(te[0])

    if rank == 0:
        # identical inputs in both arms -> identical integers
        assert w2_total == binned_total, (w2_total, binned_total)
        # size-independent property of the synthetic workload: uniform points, cutoff L/2 -> the binned fraction of all
        # ordered pairs is the sphere / cube volume ratio pi/6 (statistical scatter ~1e-6 at 1e12 pairs per frame)
        if not c5 and n_half == N_PER_TYPE:
            frac_binned = binned_total / pair_evals_total
            assert abs(frac_binned - np.pi / 6.0) < 1.0e-4, f"binned fraction {frac_binned} is not pi/6: counts are wrong"
        if c5:
            # uniform points, cutoff sphere inside the box: binned fraction = (4/3) pi r^3 / L^3 (up to the ~1e-5 of pairs
            # whose link is (anti)parallel to the axis)
            expect = 4.0 / 3.0 * np.pi * args.c5_maxdist ** 3 / L ** 3
            frac_binned = binned_total / pair_evals_total
            assert abs(frac_binned / expect - 1.0) < 2.0e-3, (frac_binned, expect)
        value = binned_total / t_value_max
        cell_evals, atomics, unique, dom_evals, dom_atomics, dom_launches, launches_all = (float(x) for x in pe[1:])
        traffic, traffic_src = _traffic_from_profiles()
        if c5:
            # pa_pdf_cells_kernel (float32 brackets for the distance AND the angle class): latency / issue bound;
            # the line carries the rate of pair evaluations after cell pruning and the staging filter, no hardware peak is claimed
            roofline = {"bound": "issue", "achieved": cell_evals / (kernel_ms_max * 1e-3) if cell_evals else None,
                        "peak": None, "unit": "pair evaluations/s", "frac": None, "traffic": None,
                        "kernel_ms_per_step": kernel_ms_max / K,
                        "note": "pa_pdf_cells_kernel: issue slots 62 % busy (ncu, 16 warps per SM at 128 registers, ~83 instructions per evaluated "
                                "pair), bins in a per-SM shared histogram; see profiles/README.md"}
        else:
            # dominant kernel: pa_rdf_cells_kernel<0,*,*,1> (float32 launch of the cell-sorted RDF kernel; 3 launches per step:
            # two symmetric same-type passes and the merged cross-type pass).  Its binding resource is the shared-memory
            # atomic unit (one ATOMS.POPC.INC per evaluated pair of a live iteration): both the atomics and the peak are
            # measured in THIS run -- the kernels count their own atomics (pa_stats.dominant_atomics) and the time of
            # exactly those launches is taken with CUDA events on their stream (pa_stats.dominant_ms).
            dom_s = dominant_ms_max * 1e-3
            roofline = {
                "bound": "smem_atomics", "kernel": "pa_rdf_cells_kernel<0,*,*,1> (float32 launch)",
                "achieved": dom_atomics / world / dom_s if dom_s else None, "peak": atomic_rate, "unit": "atomics/s",
                "frac": (dom_atomics / world / dom_s / atomic_rate) if dom_s else None,
                "useful_frac": (unique / world / (kernel_ms_max * 1e-3) / atomic_rate),
                "traffic": traffic, "traffic_source": traffic_src,
                "ms_per_launch": dominant_ms_max * world / dom_launches if dom_launches else None,
                "launches_per_step": dom_launches / world / K, "share_of_kernel_time": dominant_ms_max / kernel_ms_max,
                "atomics_per_step_per_gpu": atomics / world / K, "pair_evals_after_pruning_per_step_per_gpu": cell_evals / world / K,
                "unique_pairs_counted_per_step_per_gpu": unique / world / K,
                "kernel_ms_per_step": kernel_ms_max / K,
                "algorithmic_fp64_frac": pair_evals_rank / (kernel_ms * 1e-3) * FP64_OPS_PER_PAIR / fp64_rate,
                "fp64_issue_rate_measured": fp64_rate,
                "hbm_gbs_frame_streaming": (natoms * 12 * K) / (kernel_ms_max * 1e-3) / 1e9,
                "note": "frac = shared-memory atomics issued by the float32 launches (counted by the kernel itself) / CUDA-event "
                        "time of those launches / live-measured rate of warp-wide red.shared.add.u32 to 32 different counters "
                        "(pa_debug_smem_atomic_rate, same kernel as pa_microbench2). useful_frac = pairs that landed in a bin, "
                        "each evaluated once, / total pair-kernel time / the same peak. algorithmic_fp64_frac = SURVEY 8d's 26 "
                        "FP64-pipe instructions x ordered reference pairs / kernel time / measured DSUB-DMUL issue rate: > 1 "
                        "because cells prune, symmetric pairs are evaluated once and the class is bracketed in packed float32 "
                        "(exact fp64 only for ~1e-3 of the pairs) -- a bookkeeping figure, not a hardware fraction. HBM is not "
                        "the bound: 12 B per atom per frame.",
            }
        line = {
            "metric": METRIC if not c5 else "binned atom-pairs/s, 4M-atom 2-D distance-angle histogram (BASELINE config 5), tile shards over the GPUs",
            "value": value, "unit": "binned pairs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_value_max / K, "higher_is_better": True, "scaling": "strong" if c5 else "weak", "vs_baseline": None,
            # results are defined by the reference's fp64 arithmetic; packed float32 only brackets the class
            "dtype": "f64 (f32-bracketed class, exact f64 fix-up)" if not c5 else "f64", "data": "synthetic",
            "config": ({"workload": f"synthetic cubic box, {natoms} atoms (2 one-atom types), L={L} A, ONE frame, reference type 1 around all, "
                                    f"pdf 200x90 bins, maxdist {args.c5_maxdist} A; every GPU holds the frame and takes tile shard r of N "
                                    "(i-range split), NCCL int64 reduce once",
                        "l2": "256 MB flush between timed steps"} if c5 else
                       {"workload": f"synthetic cubic box, {natoms} atoms (2 one-atom types), L={L} A, 1 frame per GPU per step, "
                                    f"references 1 -1 and 2 -1 (all ordered pairs), maxdist_optimize, {BINS}x1 bins",
                        "pair_evals_per_step_per_gpu": pair_evals_rank / K, "frames_sharding": "frame s*N+r on GPU r, NCCL int64 reduce once",
                        "l2": "256 MB flush between timed steps; every step works on a different resident frame"}),
            "pair_evals_per_s": pair_evals_total / t_value_max,
            "e2e": {"value": w2_total / t_e2e_max, "unit": "binned pairs/s", "h2d_bytes_per_step": int(h2d // K),
                    "d2h_bytes_per_step": int(d2h // K), "ms_per_step": 1e3 * t_e2e_max / K,
                    "note": "pa_distribution_histogram_multi with host buffers: memcpy into page-locked staging, cudaMemcpyAsync, "
                            "kernels, D2H of the bins; the page-locked buffers come from the library's pool (allocated by the warm-up call)"},
            "gpu_launches": int(launches_all),
            "roofline": roofline,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            nh, fr = cpu_sample_size(threads, 15.0)
            cb = cpu_reference_run(nh, fr, threads)
            line["cpu_baseline"] = {"value": cb["binned_per_s"], "unit": "binned pairs/s", "cores": threads, "kind": cb["kind"],
                                    "pair_evals_per_s": cb["pair_evals_per_s"], "seconds": cb["seconds"],
                                    "sample": f"{cb['atoms']} atoms x {cb['frames']} frames of the same generator/density, both references "
                                              "(brute force: per-pair cost independent of N)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
