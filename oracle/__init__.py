"""CPU oracle for the cross-modal similarity hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never as the thing shipped.  The product package
(``multimodal_plankton_recognition_b200``) never imports this package.

Parity status
-------------
* Loss (``oracle.infonce``): PINNED.  The restatement is checked against the
  reference's own ``CLIPLoss`` (``/root/reference/src/coordination.py:17-47``)
  run in this container by ``oracle/make_golden.py``; the resulting vectors
  are committed under ``tests/golden/`` and re-checked by
  ``tests/test_oracle_golden.py``.
* Retrieval (``oracle.ann``): ``ANNClassifier``'s own logic
  (``/root/reference/src/ann.py:6-34``) is PINNED the same way (the reference
  class is executed with an exact index injected under the name
  ``pynndescent``).  The third-party approximate search itself
  (``pynndescent.NNDescent``, version unpinned by the reference, not installed
  in this image, no golden vectors in the reference) is restated as an exact
  euclidean search: **parity unpinned** for pynndescent's approximate graph
  search; see DESIGN.md.
"""
