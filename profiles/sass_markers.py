"""Instruction count and hardware-path mnemonics of the hot kernels, from the built library (no GPU needed):
python profiles/sass_markers.py > profiles/r2_match_sass_markers.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multibox_b200", "libmultibox_b200.so")
MARKS = ["UBLKCP", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "CREDUX", "REDUX", "MUFU.RSQ", "MUFU.LG2", "MUFU.RCP", "MUFU.EX2",
         "FFMA", "DFMA", "DADD", "DSETP", "ATOMS", "ATOM.E", "ACQBULK", "PREEXIT", "LDG.E", "STG.E", "LD.E", "ST.E",
         ".STRONG.SYS", "BAR.SYNC", "MATCH", "VOTE", "SHFL", "MEMBAR"]


def sass(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    return [ln for ln in out.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln) and ";" in ln]


def report(title, fun, first=()):
    lines = sass(fun)
    print(title)
    print("instructions: %d" % len(lines))
    for m in MARKS:
        print("%-16s %d" % (m, sum(1 for ln in lines if m in ln)))
    if first:
        print("\nfirst occurrences:")
        for m in first:
            for ln in lines:
                if m in ln:
                    print(ln.rstrip())
                    break
    print()


syms = subprocess.run(["cuobjdump", "-elf", LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
names = sorted(set(re.findall(r"_ZN3mbx\w+", syms)))
pick = lambda pat: next(n for n in names if re.search(pat, n))      # noqa: E731
print("SASS of the built libmultibox_b200.so (sm_100a; cuobjdump -sass -fun ...): instruction counts and the mnemonics that")
print("identify the hardware paths used (TMA bulk copy + mbarrier, warp reductions incl. the f32 min, fast log / rsqrt,")
print("programmatic dependent launch = PREEXIT / ACQBULK, system-scope accesses of the fused all-reduce = .STRONG.SYS).\n")
report("== mbx_match_loss_reg_kernel<8,3> (configs[1])", pick(r"mbx_match_loss_reg_kernelILi8ELi3E"),
       first=("UBLKCP", "CREDUX", "PREEXIT", "ACQBULK", ".STRONG.SYS"))
report("== mbx_match_loss_reg_kernel<8,4> (configs[3])", pick(r"mbx_match_loss_reg_kernelILi8ELi4E"))
report("== mbx_allreduce_relay_kernel (side stream: outbox -> every rank's table)", pick(r"mbx_allreduce_relay_kernel"),
       first=(".STRONG.SYS",))
report("== mbx_detect_kernel<8>", pick(r"mbx_detect_kernelILi8E"), first=("UBLKCP", "PREEXIT"))
