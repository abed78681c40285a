"""Recipe for oracle/_ref/: the UNMODIFIED reference as byte code, so that it can travel to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's CPU arms and the reference-backed tests use it; the product never
imports anything under oracle/).

The reference (rcapillon/thermoelasticity_fem) is pure Python and lives under /root/reference, which does not exist on
the GPU box.  Like a C reference that is compiled where its sources lie into ``oracle/_ref/*.so``, this script compiles the
reference's modules WHERE THEY LIE into sourceless byte code ``oracle/_ref/thermoelasticity_fem/<module>.bc`` (the .pyc format
under another extension: snapshot tools commonly drop ``*.pyc``; outputs only;
``oracle/_ref/`` is git-ignored - no reference source enters the repository or its history - but it is not gpurun-ignored, so
the built files ship with the snapshot like the CUDA library).  ``oracle/reference_shim.load_reference()`` imports the sources
when /root/reference is present and this byte code otherwise.

    python -m oracle.build_ref          # or __graft_entry__.build()
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "thermoelasticity_fem")
MODULES = ("__init__", "elements", "materials", "mesh", "model", "solvers")     # the hot path; plots / generators are not needed


def build(reference_root=None, quiet=True):
    """Returns the output directory, or None when the reference sources are not present (GPU box: use what was shipped)."""
    root = reference_root or os.environ.get("TEFEM_REFERENCE_ROOT", "/root/reference")
    src = os.path.join(root, "src", "thermoelasticity_fem")
    if not os.path.isdir(src):
        return None
    os.makedirs(OUT, exist_ok=True)
    for name in MODULES:
        py = os.path.join(src, name + ".py")
        pyc = os.path.join(OUT, name + ".bc")
        if not os.path.exists(pyc) or os.path.getmtime(pyc) < os.path.getmtime(py):
            py_compile.compile(py, cfile=pyc, dfile=f"<reference>/src/thermoelasticity_fem/{name}.py", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as fh:
        fh.write(sys.version.split()[0] + "\n")
    if not quiet:
        print("oracle/_ref built from", src)
    return OUT


if __name__ == "__main__":
    out = build(quiet=False)
    if out is None:
        print("reference sources not present: nothing built")
