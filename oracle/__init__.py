"""CPU oracle for the LSTUR hot path of nvagus/mnexp.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mnexp_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and there
only as the checker / the CPU arm, never as the product path.

PARITY UNPINNED: the reference ships no golden vectors, fixtures or
assertions for this path and its arithmetic lives in the
un-vendored, un-pinned third-party packages ``keras`` (2.2.x by idiom) and
``tensorflow`` (1.x), neither importable here.  The oracle therefore restates
the published Keras-2.2 / TF-1.x semantics and is validated by
(1) two independent implementations (float64 numpy here vs torch autograd in
``lstur_torch``), (2) hand-computed micro cases, (3) finite-difference
gradient checks and (4) structural invariants (tests/test_oracle_*.py).
"""
