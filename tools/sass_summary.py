"""Per-kernel SASS summary of libfjsp_b200.so (cuobjdump -sass): instruction count, SHA-256 of the instruction text
(tools/sass_hash.py convention) and counts of the mnemonics that prove the data path.  Runs without a GPU.

    python tools/sass_summary.py [--listing SUBSTRING_OF_MANGLED_NAME] > profiles/rNN_sass_kernels.txt
"""
import argparse
import collections
import hashlib
import os
import re
import shutil
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(REPO, "multi_agent_rl_for_fjsp_b200", "lib", "libfjsp_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UBLKCP", "UBLKPF", "SYNCS", "ACQBULK", "LDG", "ST", "LDS", "STS", "LDL", "STL", "FFMA",
        "IMAD", "LOP3", "SHFL", "VOTE", "BAR", "ATOM", "RED"]

ap = argparse.ArgumentParser()
ap.add_argument("--listing", default="fjsp_gemm_kernelILi0ELi2EE", help="print the full listing of the kernel whose mangled name contains this")
args = ap.parse_args()
tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
filt = shutil.which("cu++filt") or "/usr/local/cuda/bin/cu++filt"
out = subprocess.run([tool, "-sass", SO], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*;", ln)
        if m:
            kernels[cur].append(m.group(1))
print("SASS of libfjsp_b200.so (cuobjdump -sass, sm_100a; tools/sass_summary.py), per kernel: instruction count, SHA-256 of the")
print("instruction text (tools/sass_hash.py convention; profiles/ncu_traffic.json carries the step kernel's) and counts of the")
print("mnemonics that prove the data path: UTC*MMA / UTCBAR = tcgen05.mma / tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk")
print("(TMA bulk copies), UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier, ACQBULK = griddepcontrol; LDL / STL = local memory")
print("(spills).  Below the table: the full listing of the kernel matching --listing (default: the forward GEMM).\n")
for name, lines in kernels.items():
    dem = subprocess.run([filt, name], capture_output=True, text=True).stdout.strip() or name
    cnt = collections.Counter()
    for ins in lines:
        op = ins.split()[0] if not ins.startswith("@") else ins.split()[1]
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                cnt[k] += 1
    h = hashlib.sha256("\n".join(lines).encode()).hexdigest()
    print(dem)
    print("    %d instructions  sha256 %s" % (len(lines), h))
    print("    " + str({k: cnt[k] for k in KEYS if cnt[k]}))
for name, lines in kernels.items():
    if args.listing and args.listing in name:
        print("\n==== full listing: %s ====" % name)
        print("\n".join(lines))
        break
