"""Build-time evidence that needs no GPU: every kernel is compiled for sm_100a, the kernels on the headline lines keep their
register budget without spilling (a struct-parameter write once moved a whole argument block to local memory and cost 4 % of the
step: `profiles/r02_notes.md` §2), and the full-rank scoring kernel's SASS really contains the tcgen05 / TMA / TMEM instructions
DESIGN.md §3.2 describes (mnemonics from the B200 profiling recipe).  Reads what `cleverrec_b200.build` left in
`cleverrec_b200/build/` (objects + ptxas -v log); nothing here launches a kernel."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

BUILD = os.path.join(ROOT, "cleverrec_b200", "build")


@pytest.fixture(scope="module")
def ptxas():
    from cleverrec_b200 import build
    build.build()
    log = os.path.join(BUILD, "ptxas.log")
    if not os.path.exists(log):     # a prebuilt .so that travelled without its object directory
        build.build(force=True)
    text = open(log).read()
    info = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for '(\S+)'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\n.*?Used (\d+) registers", text):
        info[m.group(1)] = dict(arch=m.group(2), stack=int(m.group(3)), spill_st=int(m.group(4)), spill_ld=int(m.group(5)),
                                regs=int(m.group(6)))
    return info


def test_every_kernel_is_compiled_for_sm_100a_only(ptxas):
    assert len(ptxas) > 300
    assert {v["arch"] for v in ptxas.values()} == {"sm_100a"}


# (mangled-name fragment, register ceiling): the instances the bench lines and BASELINE configs run.
# bpr_step_kernel<LANES, VPL, OPT>: <32,1,*> is d = 128, <16,1,*> d = 64; OPT 0 SGD, 1 Adagrad, 3 Adam tf1.  3 CTAs x 256 threads x 80
# registers is the occupancy DESIGN.md §3.1 quotes; score_tc_kernel's 640 threads x 96 registers fill one SM's register file.
HOT = [
    ("15bpr_step_kernelILi32ELi1ELi3EE", 80), ("15bpr_step_kernelILi32ELi1ELi0EE", 64), ("15bpr_step_kernelILi32ELi1ELi1EE", 80),
    ("15bpr_step_kernelILi16ELi1ELi3EE", 80), ("15score_tc_kernel14CUtensorMap_st", 96),
    ("15loo_topk_kernelILi0EE", 80),
]


@pytest.mark.parametrize("frag,max_regs", HOT)
def test_headline_kernels_keep_their_register_budget_without_spills(ptxas, frag, max_regs):
    hits = [(k, v) for k, v in ptxas.items() if frag in k]
    assert hits, "kernel %s not found in the ptxas log" % frag
    for name, v in hits:
        assert v["spill_st"] == 0 and v["spill_ld"] == 0 and v["stack"] == 0, (name, v)
        assert v["regs"] <= max_regs, (name, v)


def test_fullrank_kernel_sass_is_tcgen05_tma_tmem():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    obj = os.path.join(BUILD, "score_tc.o")
    if not (os.path.exists(cuobjdump) and os.path.exists(obj)):
        pytest.skip("cuobjdump or the object file is not here")
    sass = subprocess.run([cuobjdump, "-sass", obj], stdout=subprocess.PIPE, text=True, check=True).stdout
    count = lambda pat: len(re.findall(pat, sass))
    assert count(r"\bUTCHMMA\b") >= 8          # tcgen05.mma (kind::f16), issued by one elected lane
    assert count(r"\bUTMALDG\.2D\b") >= 2      # cp.async.bulk.tensor.2d: the A tile and the B ring
    assert count(r"\bLDTM\b") >= 1             # tcgen05.ld: accumulators out of TMEM in the epilogue
    assert count(r"\bUTCBAR\b") >= 1           # tcgen05.commit -> mbarrier
    assert count(r"SYNCS\.ARRIVE\.TRANS64") >= 1 and count(r"SYNCS\.PHASECHK\.TRANS64") >= 1   # mbarrier expect_tx / try_wait
    assert "HMMA.16816" not in sass and "WGMMA" not in sass   # no legacy mma.sync / Hopper path behind the same name
