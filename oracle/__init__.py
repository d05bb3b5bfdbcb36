"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (plain fp32/fp64 torch functional ops, no nn.Module, no CUDA) of
the MARCONet inference hot path, used solely as the *checker* for the sm_100a
CUDA implementation in ``marconet_b200``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import anything from this package.  The product
path (``marconet_b200`` and the ``dropin/models`` mirror) never imports it and
fails loudly when the CUDA library is missing.

Parity pin: the reference ships no golden vectors or tests, and
its one third-party op (``basicsr.ops.fused_act``, un-vendored, un-pinned,
latest public release 1.4.2) is restated here from its published one-line
formula.  The restatement is pinned against *outputs of the reference's own
modules run in the build container* (``oracle/ref_harness.py`` imports
``/root/reference/models`` unmodified, ``oracle/make_golden.py`` writes the
fixtures under ``tests/golden/``).  See DESIGN.md "Oracle".
"""
