"""CPU oracle for the FinEtoolsFlexStructures.jl element-level hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import or execute it, and only as the checker or
the timed CPU baseline.  The shipped GPU path (``libfsgpu.so`` and the host
mirror in ``finetoolsflexstructures.jl_b200/``) never routes through here.

What it is: a float64 NumPy restatement of the reference algorithms
(``/root/reference/src/*.jl``, v3.6.4) for the T3FF / Q4RS shells, their
composite variants, the corotational beam, the FinEtools assemblers and the
explicit central-difference loop.  Every function cites the reference
file:line it follows.  The arithmetic that lives in un-vendored third-party
packages (FinEtools 8.2.5, FinEtoolsDeforLinear 3.0.6, SparseMatricesCSR
0.6.12, Julia 1.12 SparseArrays -- pinned in ``Manifest.toml:210-220,532-536``)
is restated in ``fe_external.py`` from its published behaviour.

Pinning status (see tests/test_oracle_goldens.py):
  * element formulations, assemblers-through-solves: PINNED by the reference's
    own known-answer tests (Scordelis-Lo T3FF/Q4RS, FV12, Nayak composite
    plate to 1e-13, Barbero 3.1 T3/Q4, layup matrices, beam buckling,
    assembler equivalence test_utilities.jl).
  * COO emission order, CSC colptr/rowval arrays, explicit-zero retention and
    the value-dependent zero dropping of SysmatAssemblerSparseSymm: PARITY
    UNPINNED by any reference test (no Julia in this image); they follow the
    documented semantics of ``SparseArrays.sparse`` / sparse ``+``.
"""
