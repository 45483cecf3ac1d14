"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (torch fp32 / numpy) restatement of the reference's denoising sampling
path (ktncktnc/diffusion-forcing-transformer): scheduling matrices, history
guidance branch tables / prepare / compose, the per-frame DDIM update and the
DiT3D (variant=full, rope_3d) backbone forward.  Every function cites the
reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
the reported CPU baseline — never on the product path.

Parity pinning: the reference ships no numerical tests, golden vectors or
fixtures for this path, so *upstream* leaves parity unpinned.
We pin the oracle ourselves against outputs of the reference executed in the
authoring container through ``oracle/ref_shim.py``: ``oracle/make_goldens.py``
writes ``tests/golden/*`` and ``tests/test_oracle_vs_golden.py`` checks every
oracle function against them (integers bit-exact, fp32 tensors <= 1e-5).
"""
