"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Molly omics-embedding hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product
path (``molly_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Parity pinning (see DESIGN.md "Oracle"):
  * stock-ESM architectures (ESM-2 rotary, NT-v1 learned-absolute): PINNED -- the
    restatement in ``esm_oracle.py`` is checked here against the reference's own
    ``OmicsOne.process_omic_sequences`` (``/root/reference/src/model/omics_one.py:49-136``)
    driving HuggingFace ``EsmForMaskedLM`` (transformers 5.5.0, eager, fp32); the
    outputs are committed as fixtures under ``tests/golden/`` by ``make_golden.py``.
  * NT-v2 gated-SiLU FFN: **parity unpinned** -- the architecture lives in un-vendored
    HF-hub remote code that is not on disk; it is restated from the public model card.
"""
