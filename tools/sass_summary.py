"""SASS instruction mix of the streaming kernels in libehmm_b200.so (cuobjdump -sass): fp64 arithmetic, special-function
unit, shared / global memory, shuffles, atomics, spills (local memory).  usage: python tools/sass_summary.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "epigrahmm_b200", "libehmm_b200.so")
WANT = ["fb_apply_kernel", "fb_reduce_kernel", "emis_inner_kernelILi6ELi6ELb1ELb1", "emis_diff_hmm_kernelILi6ELb1", "gather_consensus_kernel",
        "rc_apply_kernelILb1EjLb1ELb1ELi8", "rc_apply_kernelILb0EjLb1ELb0ELi2", "vit_scan_kernel", "glm_mix_kernel", "mt_generate_multi_kernel",
        "rs_scatter_kernel", "ps_apply_kernel"]
CLASSES = [("fp64 fma/add/mul", r"^(DFMA|DADD|DMUL)"), ("fp64 compare/select", r"^(DSETP|DMNMX)"), ("MUFU (sfu)", r"^MUFU"),
           ("int64/32 arithmetic", r"^(IMAD|IADD3|LEA|LOP3|SHF|ISETP|IABS|FLO|POPC|SEL|I2F|F2I|I2I|FSEL|PRMT|BREV|IMNMX)"),
           ("shared ld/st", r"^(LDS|STS|LDSM)"), ("global ld", r"^(LDG|LD\.)"), ("global st", r"^(STG|ST\.)"), ("atomics / red", r"^(ATOM|ATOMS|ATOMG|RED)"),
           ("shuffle / vote / match", r"^(SHFL|VOTE|MATCH)"), ("barrier", r"^(BAR|WARPSYNC|NANOSLEEP)"), ("local (spill) ld/st", r"^(LDL|STL)"),
           ("branch / exit", r"^(BRA|EXIT|CALL|RET|BSSY|BSYNC)")]
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
cur, mix = None, {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        mix[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        mix[cur][m.group(1)] += 1
print("static SASS instruction counts (sm_100a), %s" % os.path.basename(SO))
for want in WANT:
    for fn, c in mix.items():
        if want in fn:
            tot = sum(c.values())
            short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            short = short.replace("ehmm::(anonymous namespace)::", "").replace("void ", "")
            short = re.sub(r"\(.*", "", short)[:110]
            print("\n%s   [%d instructions]" % (short, tot))
            for name, pat in CLASSES:
                n = sum(v for k, v in c.items() if re.match(pat, k))
                if n:
                    print("   %-26s %6d  (%4.1f%%)" % (name, n, 100.0 * n / tot))
