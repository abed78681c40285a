"""Build compile-time variants of the assembly kernel next to the default library and print the ONE gpurun command that
A/Bs them (tools/asm_ab.py checks every pieces-first run bitwise against the first variant of its process).

    python tools/k1_variants.py            # builds thermoelasticity_fem_b200/build/variants/libtefem_<name>.so
    /usr/local/graft/bin/gpurun --timeout 300 -- '<printed command>'

Variants: name -> extra nvcc flags (see csrc/tefem_assemble.cu, csrc/tefem_element.cuh).  Run-time knobs need no build:
tools/asm_ab.py reads TEFEM_ASM_THREADS=64|128 (-> DeviceMesh.asm_threads) and takes the item layout / bank numbering / pitch
as fields of its variant strings.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermoelasticity_fem_b200 import build as B  # noqa: E402

VARIANTS = {
    "tabload": ["-DASM_TAB_LOAD"],                       # round-1 form: S_a / NN_ab loaded from shared-memory tables
    "persm768": ["-DASM_THREADS_PER_SM=768"],            # register cap 85: three 256-thread CTAs per SM where shared memory allows
}


def main():
    B.build()
    out_dir = os.path.join(B.PKG, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    others = [os.path.join(B.PKG, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != "tefem_assemble.cu"]
    libs = {}
    for name, flags in VARIANTS.items():
        obj = os.path.join(out_dir, f"asm_{name}.o")
        lib = os.path.join(out_dir, f"libtefem_{name}.so")
        subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + B.NVCC_FLAGS + flags +
                       ["-c", os.path.join(B.CSRC, "tefem_assemble.cu"), "-o", obj], check=True)
        subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-shared", "-o", lib, obj] + others + ["-ldl"], check=True)
        libs[name] = os.path.relpath(lib, ROOT)
    runs = ["echo == default and bank-aware numberings; python tools/asm_ab.py 100 512,620,8,1 512,620,8,1,1 512,620,8,1,2 512,620,8,1,3",
            "echo == 128 threads; TEFEM_ASM_THREADS=128 python tools/asm_ab.py 100 256,310,8,1 384,465,8,1",
            "echo == 64 threads; TEFEM_ASM_THREADS=64 python tools/asm_ab.py 100 128,160,8,1 192,235,8,1 256,310,8,1"]
    for name, lib in libs.items():
        runs.append(f"echo == {name}; TEFEM_LIB=$PWD/{lib} python tools/asm_ab.py 100 512,620,8,1")
    print("mkdir -p gpurun_out; (" + "; ".join(runs) + ") > gpurun_out/k1_variants.log 2>&1; cat gpurun_out/k1_variants.log")


if __name__ == "__main__":
    main()
