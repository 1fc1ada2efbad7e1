"""Offline (no GPU) compile of the tile kernel for one example problem with nvcc for sm_100a:
registers / spills / shared memory from `-Xptxas -v`, and optionally the SASS of the kernel.

    python tools/sass_check.py delta_iii [--flags 12] [--threads 128] [--min-blocks 3] [--sass out.sass]
                               [-D NAME[=V] ...]
"""
import argparse, os, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from examples import problems
from examples.cases import lower_case

NAMES = {"delta_iii": "delta_iii_launch_vehicle", "cartpole": "cart_pole_swing_up"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("problem")
    ap.add_argument("--flags", type=int, default=12)
    ap.add_argument("--threads", type=int, default=128)
    ap.add_argument("--min-blocks", type=int, default=3)
    ap.add_argument("--sass", default=None)
    ap.add_argument("-D", action="append", default=[])
    a = ap.parse_args()
    fn = getattr(problems, NAMES.get(a.problem, a.problem))
    low, _, _ = lower_case(fn(), "lobatto", 40, 4, seed=0)
    csrc = os.path.join(ROOT, "pycollo_b200", "csrc")
    d = tempfile.mkdtemp(prefix="pcx_sass_")
    open(os.path.join(d, "pcx_problem.h"), "w").write(low.header)
    shutil.copy(os.path.join(csrc, "pcx_kernels.cuh"), os.path.join(d, "k.cu"))
    cub = os.path.join(d, "k.cubin")
    cmd = ["nvcc", "-O3", "-std=c++17", "-cubin", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
           f"-I{d}", f"-I{csrc}", f"-DPCX_FLAGS={a.flags}", f"-DPCX_THREADS={a.threads}",
           f"-DPCX_MIN_BLOCKS={a.min_blocks}", f"-DPCX_KERNEL_NAME=pcx_fill_{a.flags}", "-diag-suppress=177",
           "-Xptxas", "-v", *[f"-D{x}" for x in a.D], "-o", cub, os.path.join(d, "k.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    print((r.stdout + r.stderr).strip())
    if r.returncode:
        sys.exit(1)
    if a.sass:
        s = subprocess.run(["cuobjdump", "-sass", cub], capture_output=True, text=True).stdout
        open(a.sass, "w").write(s)
        print("SASS lines:", s.count("\n"), "->", a.sass)


if __name__ == "__main__":
    main()
