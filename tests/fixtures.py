"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the
reference's own binary) -> capi.Trajectory + job requests + expected outputs."""
import os

import numpy as np

from prealpha_b200 import capi, inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Fixture:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.types = [tuple(int(v) for v in row) for row in z["types"]]
        self.nsteps = int(z["nsteps"])
        coords = z["coords"]
        if "replicate" in z.files:
            coords = np.repeat(coords, int(z["replicate"]), axis=0)
        self.coords = coords                      # [nsteps, natoms_total, 3]
        self.elements = [str(e) for e in z["elements"]]
        self.box_lo, self.box_hi = z["box_lo"], z["box_hi"]
        self.counts = [tuple(int(v) for v in r) for r in z["counts"]]
        self.job_names = [str(s) for s in z["job_names"]]
        self.job_texts = [str(s) for s in z["job_texts"]]
        self.files = {str(fn): z[f"file_{i}"].tobytes() for i, fn in enumerate(z["file_names"])}
        self.stdout = z["stdout"].tobytes().decode()
        self.atom_charges = [z[f"atom_charge_{i}"] for i in range(len(self.types))] if "atom_charge_0" in z.files else None
        self.lmp = "lmp" in z.files
        self.wrap = "wrap" in z.files
        self.unwrap = "unwrap" in z.files
        self.mol_tail = str(z["mol_tail"]) if "mol_tail" in z.files else ""            # extra sections of molecular.inp
        self.general_extra = str(z["general_extra"]) if "general_extra" in z.files else ""

    def trajectory(self) -> capi.Trajectory:
        tl, off = [], 0
        for ti, (ch, na, nm) in enumerate(self.types):
            n = na * nm
            xyz = self.coords[:, off:off + n, :].reshape(self.nsteps, nm, na, 3)
            d = dict(charge=ch, natoms=na, nmol=nm, elements=self.elements[off:off + na], xyz=xyz)
            if self.atom_charges is not None:
                d["atom_charge"] = self.atom_charges[ti]
            tl.append(d)
            off += n
        traj = capi.Trajectory(tl, self.box_lo, self.box_hi, xyz_format=not self.lmp, wrap_trajectory=self.wrap)
        if self.wrap or self.unwrap:
            # the reference rewrites the stored float32 coordinates before any analysis (Molecular:4801-4816);
            # tests apply the oracle's restatement of those pre-passes to the fixture's raw coordinates
            from oracle import oracle_lib
            (oracle_lib.wrap_full if self.wrap else oracle_lib.unwrap_full)(traj)
        return traj

    def jobs(self):
        """[(job_name, reference_number (1-based within its input file), pa_job_request, (within, oob))]"""
        out, k = [], 0
        for jn, text in zip(self.job_names, self.job_texts):
            for n, rq in enumerate(inputs.parse_distribution_input(text, len(self.types)), start=1):
                out.append((jn, n, rq, self.counts[k]))
                k += 1
        return out

    def expected_files(self, job_name):
        pre = job_name + "_"
        return {fn[len(pre):]: c for fn, c in self.files.items() if fn.startswith(pre)}


ALL = ["bmim", "re4", "re4_atoms", "synth", "lmp", "lmp_wrap", "lmp_unwrap", "charge_arm"]


def available():
    return [n for n in ALL if os.path.exists(os.path.join(GOLDEN, n + ".npz"))]
