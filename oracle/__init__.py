"""CPU oracle for the federated hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, in plain PyTorch-CPU / numpy fp32, the arithmetic of the
reference's client-training -> update-level DP -> FedAvg path so the CUDA kernels
can be checked against it.  Every function cites the reference file:line it
follows (paths relative to the upstream repository root).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``federated-learning-for-privacy-preserving-image-classification_b200/`` (the
product) imports it, and the product has no CPU fallback.

Pinning: the reference ships no golden vectors (SURVEY.md "Five facts" #4), so
the oracle is pinned against the reference ITSELF: ``oracle/make_golden.py``
imports the unmodified reference leaf modules from ``/root/reference`` (possible
only in the build container), runs them on seeded inputs and stores inputs +
outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays the
oracle on those inputs.  The per-sample DP-SGD oracle (``oracle/dpsgd.py``) has
no reference implementation at all -- it is "parity unpinned" and says so.
"""
