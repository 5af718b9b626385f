"""TEST INFRASTRUCTURE ONLY -- builds tests/emu/_build/libgfmd_b200_emu.so: the kernel and host
sources of user-gfmd_b200/csrc compiled for the CPU against the emulation shim
(tests/emu/include/cuda_runtime.h).  Used by tests/test_emulated_kernels.py; the product never
loads it."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "user-gfmd_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libgfmd_b200_emu.so")
FAKE_NCCL = os.path.join(OUT, "libgfmd_fake_nccl.so")

sys.path.insert(0, HERE)
import preprocess  # noqa: E402


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")))


def up_to_date():
    if not os.path.exists(LIB) or not os.path.exists(FAKE_NCCL):
        return False
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(HERE, "*.py")) + glob.glob(os.path.join(HERE, "*.cpp")) + \
        glob.glob(os.path.join(HERE, "include", "*.h")) + [os.path.join(ROOT, "include", "gfmd_b200.h")]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False):
    if not force and up_to_date():
        return LIB
    src = os.path.join(OUT, "src", "user-gfmd_b200", "csrc")
    os.makedirs(src, exist_ok=True)
    os.makedirs(os.path.join(OUT, "src", "include"), exist_ok=True)
    for p in sources():
        with open(p) as f:
            text = preprocess.rewrite(f.read())
        name = os.path.basename(p)
        if name.endswith(".cu"):
            name = name[:-3] + ".cpp"
        with open(os.path.join(src, name), "w") as f:
            f.write(text)
    # the C ABI header is included by relative path ("../../include/gfmd_b200.h")
    with open(os.path.join(ROOT, "include", "gfmd_b200.h")) as f:
        hdr = f.read()
    with open(os.path.join(OUT, "src", "include", "gfmd_b200.h"), "w") as f:
        f.write(hdr)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-march=native", "-ffp-contract=off", "-fPIC", "-shared", "-w",
           "-I", os.path.join(HERE, "include"), "-o", LIB,
           os.path.join(src, "gfmd_b200.cpp"), os.path.join(HERE, "emu_runtime.cpp"), "-ldl"]
    subprocess.check_call(cmd)
    # in-process stand-in for NCCL (ranks = host threads), see fake_nccl.cpp
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-I", os.path.join(HERE, "include"),
                           "-o", FAKE_NCCL, os.path.join(HERE, "fake_nccl.cpp"), "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
