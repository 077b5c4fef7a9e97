"""Debug helper: the end-to-end C-ABI call of bench.py on 4 frames of the 1M-atom workload, with and without the float32
launch (PA_EXEC_NO_F32 = 8); prints wall time, kernel time and counters.  PA_DEBUG_TIMING=1 adds the host phases."""
import time, sys, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from prealpha_b200 import api
n_half=bench.N_PER_TYPE; natoms=2*n_half; L=float(bench.L_BOX)
frames=np.stack([bench.synth_frame(s, natoms, L) for s in range(4)])
traj=bench.make_traj(frames, n_half, L); jobs=bench.make_jobs(traj)
for flags in (0, 0, 8, 8):
    t0=time.perf_counter()
    h,w,o,st=api.histogram_multi(traj, jobs, api.make_exec(device=0, frames_per_batch=1, flags=flags))
    t=time.perf_counter()-t0
    print("flags", flags, "wall %.3f s"%t, "kernel_ms %.1f"%st.kernel_ms, "launches", st.kernel_launches, {n: getattr(st,n) for n,_ in st._fields_ if n not in ("kernel_ms","kernel_launches")})
