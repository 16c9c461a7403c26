"""CPU oracle for the FD-GAN hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``fdgan_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` use it, and only as the checker / the CPU arm being timed.

Parity status: the reference ships no golden vectors, so the oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, generated in the authoring container
by ``oracle/make_golden.py`` (imports ``/root/reference`` through the shims in
``oracle/ref_import.py``) and committed under ``tests/golden/``.  The one piece
that cannot be executed (``loss.py`` survives only as CPython-3.6 bytecode,
SURVEY Appendix B) is restated from the disassembly and is "parity unpinned".
"""
