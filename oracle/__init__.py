"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SemiBin clustering front end.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product (``semibin_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * kNN restatement   : pinned against live scikit-learn 1.9.0 (the version the
                        reference locks, pixi.lock:4152-4153) on the reference's
                        own fixture test/bin_data and on synthetic tie/duplicate
                        sets -> tests/golden/*.npz  (oracle/make_golden.py).
  * cal_kl restatement: pinned against SemiBin.cluster.cal_kl (numpy evaluator)
                        and the reference's own slow_cal_kl (test/test_cal_kl.py).
  * edge list         : pinned against the UNMODIFIED run_embed_infomap executed
                        with a stub ``igraph`` module (hand-off #1 capture).
  * DBSCAN sweep      : pinned against live sklearn.cluster.dbscan on the
                        unmodified call of long_read_cluster.py:115-120.
"""
