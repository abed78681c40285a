"""CPU oracle for the thermoelastic assembly + Dirichlet + solve hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker / the CPU arm that is timed next to the GPU path.  The product package
``thermoelasticity_fem_b200`` never imports it and has no CPU fallback.

Parity pinning: the reference (rcapillon/thermoelasticity_fem) ships no tests
and no golden vectors.  The restatement in
``oracle/tefem_oracle.py`` is therefore pinned against outputs of the
reference itself, produced in the build container by importing the reference's
own modules through ``oracle/reference_shim.py`` and committed as fixtures in
``tests/golden/`` by ``oracle/make_golden.py``.
"""
