"""CPU oracle for the RANSAC-Flow per-pair inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline - never on the CUDA product path (``ransac-flow_b200/`` never
imports this package and fails loudly when its CUDA library is missing).

What it is: a restatement, in numpy / torch-CPU fp32, of the reference's
algorithm for the hot path, each function
citing the reference ``file:line`` it follows (paths relative to the
reference checkout, ``/root/reference`` in the build container).

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the reference itself:
``oracle/gen_golden.py`` imports the unmodified reference modules
(``utils/outil.py``, ``model/model.py``, torchvision ResNet-50) in the build
container and stores their outputs for seeded inputs in ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every oracle function against them.
One piece stays *unpinned*: ``warp_grid`` restates kornia==0.1.4.post2
``HomographyWarper.warp_grid`` (requirements.txt:59), a third-party
dependency that is not vendored in the reference and not installed here.
"""
