"""SHA-256 of the SASS of one kernel of libfjsp_b200.so (instruction text only: addresses and encodings stripped), so a
measured ncu figure can be tied to the binary it was measured on (profiles/ncu_traffic.json, bench.py roofline.traffic).

    python tools/sass_hash.py [substring of the demangled kernel name]   # default: the step kernel the bench times
"""
import hashlib
import os
import re
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(REPO, "multi_agent_rl_for_fjsp_b200", "lib", "libfjsp_b200.so")
STEP_KERNEL = "fjsp_step_kernelILi1ELb0ELb0EE"  # fjsp::fjsp_step_kernel<1, false, false>: 1 cell, float outputs, compact layout


def kernel_sass(pattern=STEP_KERNEL, so=SO):
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([tool, "-sass", so], capture_output=True, text=True, check=True).stdout
    lines, take = [], False
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            take = pattern in m.group(1)
            continue
        if take:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*;", ln)
            if m:
                lines.append(m.group(1))
    return lines


def sass_sha256(pattern=STEP_KERNEL, so=SO):
    lines = kernel_sass(pattern, so)
    if not lines:
        return None, 0
    return hashlib.sha256("\n".join(lines).encode()).hexdigest(), len(lines)


if __name__ == "__main__":
    h, n = sass_sha256(sys.argv[1] if len(sys.argv) > 1 else STEP_KERNEL)
    print(h, n)
