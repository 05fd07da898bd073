"""CPU oracle for the SVD-dagger world-model hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``world-in-world_b200/`` (the product)
may import, call, link or execute anything in this package.  The only allowed
importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the *checker* or the
*baseline*, never as the thing measured or shipped.

Pinning: ``oracle/make_golden.py`` imports the reference's vendored diffusers
fork (``/root/reference/FTsvd/diffusers-private``, diffusers 0.31.0) in the build
container, runs it on seeded inputs and freezes its outputs under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures, so parity is pinned to outputs of the reference itself.
Unpinned boundaries (third-party arithmetic not in /root/reference): the CLIP
image encoder (``transformers``) and PIL bicubic resize -- see DESIGN.md.
"""
