"""A few cases of every differential fuzzer of tests/golden/ in the CPU suite: the drop-in's host logic (input readers,
job set-up, bookkeeping of the cluster / speciation replays, normalisation, writers, report formats) runs unchanged with
its GPU entry point interposed by the oracle's loops and must write byte-identical files / report blocks to the reference
binary's on random inputs.  Needs the reference binary (oracle/_ref, staged by build() where /root/reference exists);
skipped elsewhere.  The scripts take (ncases, seed) and end with `mismatches: N of M`."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,ncases,seed", [("fuzz_distribution.py", 6, 101), ("fuzz_streaming.py", 5, 105), ("fuzz_distance.py", 6, 102), ("fuzz_cluster.py", 6, 103),
                                                ("fuzz_speciation.py", 6, 104), ("fuzz_general.py", 6, 106)])
def test_host_logic_equals_the_reference_binary_on_random_inputs(script, ncases, seed, oracle):
    if not oracle.ref_available():
        pytest.skip("reference binary not staged (oracle/_ref)")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", script), str(ncases), str(seed)], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    tail = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
    m = re.match(r"mismatches: (\d+) of (\d+)", tail)
    assert p.returncode == 0 and m, p.stdout[-2000:] + p.stderr[-2000:]
    assert int(m.group(1)) == 0, p.stdout[-4000:]
    compared = sum(1 for ln in p.stdout.splitlines() if ": ok" in ln)
    assert compared >= ncases // 2, p.stdout[-2000:]            # (the binary itself fails on some odd inputs: those are skipped)
