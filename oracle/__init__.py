"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the reference's region-captioning path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as
the checker / the timed CPU baseline.  The product package (``grasp_any_region_b200``)
never imports ``oracle`` and raises if its CUDA library is missing.

Parity pinning status:
  * GAR glue (mask-id decode, mask conv, masked_scatter, RoI replay), projector, Llama,
    greedy loop: PINNED against the reference's own modules executed in this container
    (``oracle/ref_harness.py`` -> ``tests/golden/ref_tiny_*.npz``).
  * PE ViT (timm Eva): the reference does not vendor timm and timm is not installed here,
    so the vision tower is a restatement from the upstream algorithm.  PARITY UNPINNED for
    that one module (the reference's call protocol around it -- modeling_perception_lm.py
    :181-217 -- is exercised unmodified by the harness).
"""
