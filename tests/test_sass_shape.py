"""Build-time checks on the compiled traversal kernels (CPU; reads the SASS of libbpt.so with cuobjdump).

The scheduling loop of persistent_trace (csrc/trace.cuh) votes with ONE warp-wide REDUX per iteration and every lane must
reach that same instruction.  A build in which nvcc/ptxas rotated and peeled the loop -- two copies of the vote -- hung on
B200 (see the comment on the loop); the source keeps the loop in shape with a bounded trip count.  This test pins the
compiled shape: exactly one REDUX.SUM per traversal kernel, (next to) no local-memory spills in the hot kernels, and the register
budget that the occupancy in kernels.cuh (BPT_TRACE_MIN_CTAS CTAs x 128 threads) relies on."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "buas_pathtracer_b200", "libbpt.so")
PTXAS_LOG = os.path.join(ROOT, "buas_pathtracer_b200", "csrc", "build", "ptxas.log")


def _sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True).stdout
    kernels = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[name].append(line)
    return kernels


def test_traversal_kernels_vote_once_per_iteration(bpt):
    kernels = _sass()
    hot = {k: v for k, v in kernels.items() if re.search(r"k_trace_(closest|shadow|merged)|k_tail|k_trace_api|k_recursive", k)}
    assert len(hot) >= 8, sorted(kernels)
    for name, lines in hot.items():
        redux = [l for l in lines if "REDUX.SUM" in l]
        # flush_ray_counts / flush_counters add their own REDUX after the loop in some kernels; the loop's vote is the one
        # that feeds a uniform register (UR) compared against zero
        votes = [l for l in redux if re.search(r"REDUX\.SUM\s+UR", l)]
        assert len(votes) >= 1, name
        flushes = 2 if re.search(r"k_tail|k_recursive", name) else 0           # flush_ray_counts after the loop
        loop_votes = len(votes) - flushes
        if "k_recursive" in name:
            # ptxas tail-duplicates the loop head of this (much larger) kernel into its phase branches: four copies
            # of the vote, each reached by the whole warp (the branch conditions are warp-uniform).  It has run clean on
            # every test; the count is pinned so that a change of shape is noticed here first.
            assert loop_votes == 4, f"{name}: {loop_votes} copies of the loop vote (4 expected)"
        else:
            assert loop_votes == 1, f"{name}: {loop_votes} copies of the loop vote -- the scheduling loop was duplicated (peeled / rotated)?"


def test_hot_traversal_kernels_do_not_spill(bpt):
    if not os.path.exists(PTXAS_LOG):
        pytest.skip("no ptxas log (library built elsewhere)")
    text = open(PTXAS_LOG).read()
    blocks = re.split(r"ptxas info\s+: Compiling entry function '", text)[1:]
    seen = 0
    for b in blocks:
        name = b.split("'")[0]
        if not re.search(r"k_trace_(closest|shadow)ILb0|k_trace_merged", name):
            continue
        seen += 1
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        regs = re.search(r"Used (\d+) registers", b)
        assert spill and regs, name
        assert int(spill.group(1)) <= 16 and int(spill.group(2)) <= 16, f"{name} spills: {spill.group(0)}"     # a register or two at most
        assert int(regs.group(1)) <= 56, f"{name}: {regs.group(1)} registers do not fit 9 CTAs x 128 threads per SM"
    assert seen == 3
