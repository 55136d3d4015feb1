"""CPU oracle for the BNN hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch-CPU / numpy) restatement of the reference's
algorithm for the posterior-sample-batched MLP forward and its predictive
reduction.  It exists so that the CUDA path can be checked against something.

Rules (see DESIGN.md "Oracle"):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it;
  * nothing under ``bnn_b200/`` imports it, and the product path has no CPU
    fallback -- it raises if the CUDA library is missing;
  * parity status: PINNED for the predictive path, the BBB draw + KL, the
    MC-dropout and concrete-dropout draws (against fixtures generated from the
    unmodified reference, ``tests/golden/make_golden.py``); UNPINNED for the
    pSGLD update (``pybnn`` is absent offline) and the SVI guide draw (``pyro``
    is absent) -- both restate the published algorithm only.
"""
