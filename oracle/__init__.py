"""CPU oracle for the LSTUR hot path of nvagus/mnexp.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mnexp_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and there
only as the checker / the CPU arm, never as the product path.

PARITY: PINNED ON THE REFERENCE'S OWN CODE, KERAS RESTATED.  The reference
ships no golden vectors, fixtures or assertions for this path (SURVEY.md §4,
§8c) and its layer arithmetic lives in the un-vendored, un-pinned third-party
packages ``keras`` (2.2.x by idiom) and ``tensorflow`` (1.x), neither
importable here.  So:

* ``oracle/keras_shim/`` restates the published behaviour of the Keras / TF
  calls the path makes (its README.md lists them), and with it on sys.path the
  reference's own unmodified ``task/paper.py``, ``task/cook.py``,
  ``task/seq2vec.py``, ``models.py``, ``document.py``, ``utils.py``,
  ``settings.py`` are IMPORTED FROM /root/reference AND RUN
  (``tests/golden/make_ref_golden.py``): its loaders, batchers and
  ``_build_model`` graphs produce ``tests/golden/ref_golden.npz`` — 44 task
  class / user encoder / scorer cases with batches, forward outputs, losses,
  gradients and Adam steps, the decomposed pipeline, and five whole runs of the
  reference's ``main.train`` / ``main.cook`` command functions with everything
  they log.
* ``lstur_numpy`` (float64) and ``lstur_torch`` (autograd + Keras Adam), written
  independently a round earlier from a reading of the reference, reproduce
  those vectors to 1e-9 (``tests/test_ref_pinned.py``); where they did not —
  cook 'atgru' pools 2U one-feature steps, not two U-wide ones — the oracle
  and the CUDA path were corrected to what the reference's code computes.
* What remains restated rather than executed is the arithmetic INSIDE each
  Keras layer / backend function (GRU cell, Conv1D 'same', Masking, Adam,
  categorical_crossentropy ...): two independent restatements (the shim's
  layers, the oracle's closed forms) that agree, cross-checked by hand-computed
  micro cases and finite differences (tests/test_oracle.py).
"""
