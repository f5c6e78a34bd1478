"""oracle/ — TEST INFRASTRUCTURE ONLY (never imported by the product package).

A CPU (PyTorch fp32/fp64) restatement of the reference's algorithm for the hot
path: momentum-PGD, training APGD, evaluation APGD, the FARE/TeCoA losses, and
the open_clip-style CLIP vision transformer they wrap.  Every function cites the
reference file:line it follows (paths relative to chs20/RobustVLM).

Who may import this package: `tests/`, `__graft_entry__.smoke()`, and
`bench.py`'s `cpu_baseline` / `--impl reference` legs — as the checker or the
timed CPU baseline, never as the product.

Parity pinning
--------------
* Attack / loss code (`oracle.attacks`): pinned against the *real* reference
  functions imported from /root/reference in the authoring container
  (`tests/test_oracle_vs_reference.py`, bit-exact on CPU) and against golden
  vectors generated from the reference by `tests/golden/make_golden.py`
  (committed under `tests/golden/`).
* ViT (`oracle.vit`): the reference does not contain the model — it lives in
  the un-vendored pip dependency `open-clip-torch==2.19.0`
  (requirements.txt:91), absent here and un-installable (no network).  The
  restatement follows open_clip 2.19.0's published `VisionTransformer`
  semantics and is cross-checked against the
  independent HF `transformers.CLIPVisionModelWithProjection`
  (`tests/test_oracle_vit.py`).  The reference itself holds no golden vectors
  or tests for this path, so for the ViT: **parity unpinned**
  beyond that cross-check.
"""
