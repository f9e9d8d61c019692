"""TEST INFRASTRUCTURE -- not product code.

CPU/torch restatement ("oracle") of the hot path of the KUIS-AI LHBDC / Flex-Rate
B-frame codecs: flow-driven bilinear backward warp, GDN/IGDN, and the
quantise + Gaussian-conditional / factorised-hyperprior likelihood.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``video-compression_b200/b200vc``) never imports it and has no CPU fallback.

Parity status
-------------
* warp / glue / blend / bit sums: the arithmetic lives in the reference itself
  (``LHBDC/model/m.py``, ``LHBDC/model/flow.py``, ``Flex-Rate.../b_model/b_model.py``,
  ``ICIP2024/src/model/m.py``) and in torch's ``grid_sample``.  Pinned: the
  restatement is checked against the reference modules imported verbatim in the
  build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
* GDN / EntropyBottleneck / GaussianConditional: the arithmetic lives in the
  third-party dependency ``compressai==1.1.8`` (``LHBDC/environment.yml:142``),
  which is absent from ``/root/reference`` and not installable offline.  The
  restatement in ``oracle/cai.py`` follows CompressAI 1.1.x's published algorithm
  and the reference call sites (``LHBDC/model/layers.py:6-17,93-117``).
  **parity unpinned** at the CompressAI boundary: the reference ships no golden
  vector or test for it; we pin what can be pinned (closed-form identities and
  the reference's own model files running *through* the restated ops).
"""
