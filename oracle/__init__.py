"""CPU oracle for the LightGCN hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as
the checker or as the timed CPU baseline.  The product package never imports it and fails
loudly when its CUDA library is missing.

Parity status
-------------
* ``reference_path.py`` restates the reference's OWN lines (models/light_gcn.py:28-40,
  utils/helpers.py:84-102, utils/train_test.py:18-64,82-103,105-134,165-212,
  utils/recommend.py:35-61, data/dataset_handler.py:273-282).  It is PINNED: the fixtures in
  ``tests/golden/`` were produced by importing the UNMODIFIED reference modules from
  ``/root/reference`` (script: ``oracle/gen_golden.py``) and the restatement is checked
  against them in ``tests/test_oracle_golden.py``.
* ``pyg_restated.py`` restates third-party code the reference calls but does not vendor
  (torch-geometric==2.4.0, environment.yml:24; pytorch-sparse -> METIS 5.1.0, README.md:31).
  Those packages are not installable here, so that part is **parity unpinned**: it follows
  the published PyG 2.4.0 algorithm and is anchored by closed-form known answers
  (models/light_gcn.py:66-89 smoke graph) and an independent torch.sparse CSR cross-check.
"""
