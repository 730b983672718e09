"""profiles/r02_ptxas.md: registers / stack / spills / static shared memory of every kernel, from `nvcc -Xptxas -v` of each
.cu file with the library's flags (no GPU needed).  python tools/ptxas_summary.py [out.md]"""
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tdnn-f_nas_b200"))
import build as B  # noqa: E402


def one(src, tmp):
    r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-Xptxas", "-v", "-c", src, "-o", os.path.join(tmp, os.path.basename(src) + ".o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return os.path.basename(src), r.stderr


def short(mangled):
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"^void ", "", name.replace("(anonymous namespace)::", "").replace("tdnnf::", ""))
    depth = 0
    for i, ch in enumerate(name):
        depth += (ch == "<") - (ch == ">")
        if ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_ptxas.md")
    rows = []
    with tempfile.TemporaryDirectory() as tmp, ThreadPoolExecutor(8) as ex:
        for unit, log in ex.map(lambda s: one(os.path.join(B.CSRC, s), tmp), B.CUDA_SOURCES):
            for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes "
                                 r"spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", log, re.S):
                rows.append((unit, short(m.group(1)), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(7) or 0)))
    spills = [r for r in rows if r[4]]
    out = ["# r02: ptxas -v resource usage of every kernel in libtdnnf_nas_b200.so (sm_100a, nvcc 12.9, -O3; static shared memory only: "
           "the GEMM / slice kernels take theirs dynamically)", "",
           f"{len(rows)} kernels, {len(spills)} with register spills"
           + (" (" + ", ".join(f"`{r[1]}` {r[4]} B" for r in spills) + ")" if spills else "")
           + "; stack frames are local arrays (per-offset tables passed by value), not spills.", "",
           "| source | kernel | registers | stack B | spill stores B | static smem B |", "|---|---|---:|---:|---:|---:|"]
    for r in sorted(rows, key=lambda r: (r[0], -r[2])):
        out.append(f"| `{r[0]}` | `{r[1]}` | {r[2]} | {r[3]} | {r[4]} | {r[5]} |")
    open(out_path, "w").write("\n".join(out) + "\n")
    print(out_path, len(rows), "kernels")


if __name__ == "__main__":
    main()
