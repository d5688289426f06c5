"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithms for the hot path (see each file's header for the
reference file:line it follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The product
(``seismic_bpmf_b200``) never imports it and fails loudly when its CUDA library is missing.
"""
