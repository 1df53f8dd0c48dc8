"""CPU oracle for the OpenStereo cost-volume hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference's algorithm for the path
named by BASELINE.json:north_star (cost-volume construction -> 3D aggregation
-> soft-argmin).  It exists to CHECK the CUDA product in ``openstereo_b200``;
nothing under ``openstereo_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs execute it.

Arithmetic provenance
---------------------
The reference (XiandaGuo/OpenStereo @ 23d71c9) is pure Python on top of
PyTorch: every number it produces comes out of an aten CPU kernel (third-party
dependency ``torch``; the reference asks for ``pytorch>=1.13.1``,
docs/0.get_started.md:10; this image pins torch 2.11.0+cu128).  The oracle
therefore issues the same aten calls in the same order on fp32 CPU tensors, so
that it is BIT-EQUAL to the reference functions it restates -- that equality is
asserted against the imported reference by ``tools/make_golden.py`` (which
writes ``tests/golden/*.npz``) and by ``tests/test_oracle_pins_reference.py``
whenever ``/root/reference`` is present.

Pinning status
--------------
The reference ships NO test, golden vector or known-answer fixture for this
path ("parity unpinned" by the reference's own tests).
The oracle is pinned instead against OUTPUTS OF THE REFERENCE ITSELF, generated
in the authoring container by importing the reference modules
(``oracle/_reference_shim.py``) and committed as ``tests/golden/*.npz`` together
with the generating script ``tools/make_golden.py``.
"""
