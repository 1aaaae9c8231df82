"""CPU: what the compiler made of the hot kernels (no GPU needed: nvcc cross-compiles sm_100a).

The stage kernels' budgets -- stage_strip: 128 registers x 4 CTAs of 128 threads per SM; stage_tma (persistent, cp.async
staging): 168 registers x 3 CTAs -- with no local-memory traffic in the variants the benchmarks run;
a change that makes ptxas spill there costs more than any micro-optimisation gains.  Reads the ptxas logs that
mara3_b200/csrc/Makefile keeps and the cubin inside the library."""
import os
import re
import shutil
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "build", "csrc", "ptxas_kernels.log")
LOG_TMA = os.path.join(ROOT, "build", "csrc", "ptxas_stage_tma.log")
LIB = os.path.join(ROOT, "mara3_b200", "libmara3_b200.so")


def kernel_stats(log=LOG):
    """{mangled kernel name: (registers, spill store bytes, spill load bytes, stack bytes)} from `ptxas -v`."""
    stats, name, frame = {}, None, (0, 0, 0)
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            name, frame = m.group(1), (0, 0, 0)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            frame = (int(m.group(2)), int(m.group(3)), int(m.group(1)))
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            stats[name] = (int(m.group(1)),) + frame
            name = None
    return stats


@pytest.mark.skipif(not os.path.exists(LOG), reason="no ptxas log (the library was not built in this tree)")
def test_benchmark_variants_of_the_stage_kernel_do_not_spill():
    stats = kernel_stats()
    # stage_strip<4, 64, FAST, MODE 1 / 2, JUMP 0 / 1, QMODE 0>: the variants of configs 2-5
    hot = [k for k in stats if re.search(r"stage_stripILi4ELi64ELb1ELi[12]ELb[01]ELb0E", k)]
    assert len(hot) == 4, sorted(stats)
    for k in hot:
        regs, st, ld, stack = stats[k]
        assert regs <= 128 and st == 0 and ld == 0 and stack == 0, (k, stats[k])
    # every variant keeps the 4-CTAs-per-SM register budget
    for k, (regs, *_rest) in stats.items():
        if "stage_strip" in k:
            assert regs <= 128, (k, regs)


@pytest.mark.skipif(not os.path.exists(LOG_TMA), reason="no ptxas log (the library was not built in this tree)")
def test_persistent_stage_kernel_keeps_three_ctas_per_sm_without_spills():
    stats = kernel_stats(LOG_TMA)
    hot = [k for k in stats if re.search(r"stage_tmaILi3ELi64ELb1ELi[12]E", k)]
    assert len(hot) == 2, sorted(stats)
    for k in hot:
        regs, st, ld, stack = stats[k]
        assert regs <= 168 and st == 0 and ld == 0 and stack == 0, (k, stats[k])
    for k, (regs, *_rest) in stats.items():
        assert regs <= 168, (k, regs)


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_stage_kernel_stages_its_tiles_with_async_copies():
    """north_star: "each block's tile plus guard zones staged into shared memory by TMA or cp.async" -- cp.async is LDGSTS in SASS."""
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "stage_tma", LIB], capture_output=True, text=True, timeout=300).stdout
    if "LDGSTS" not in sass:       # (older cuobjdump: -fun wants the mangled name)
        sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=300).stdout
        sass = "".join(part for part in sass.split("Function : ") if "stage_tma" in part.split("\n", 1)[0])
    assert sass.count("LDGSTS") >= 9, sass.count("LDGSTS")


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_library_carries_sm_100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
