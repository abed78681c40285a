"""Host-emulation build of the K1 per-thread routines (TEST INFRASTRUCTURE, never loaded by the package).

Compiles thermoelasticity_fem_b200/csrc/tefem_assemble.cu with g++ and -DTEFEM_HOST_EMU so that the
SAME per-thread routines the CUDA kernels call (stage_element, assemble_block, element_dense) run
sequentially on the CPU.  It lets the `-m "not gpu"` tests check the slot-map / schedule logic of the
symbolic phase against the oracle in a container without a GPU.  The product has no CPU path.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "thermoelasticity_fem_b200", "csrc")
LIB = os.path.join(HERE, "libtefem_emu.so")


def build():
    srcs = [os.path.join(CSRC, "tefem_assemble.cu"), os.path.join(CSRC, "tefem_core.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) < os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTEFEM_HOST_EMU", "-ffp-contract=off",
           "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I/usr/local/cuda/include"]
    for s in srcs:
        cmd += ["-x", "c++", s]
    cmd += ["-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build())
