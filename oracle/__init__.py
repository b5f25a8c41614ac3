"""CPU oracle for the HSG clustering + contrastive hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the reference
algorithm (twke18/HSG, Python/PyTorch) for the path named in BASELINE.json.  It
is imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- always as the
checker or as the timed CPU baseline, never as part of the product path.  The
product (``hsg_b200``) never imports it and fails loudly when its CUDA library
is missing.

Parity status: PINNED.  The reference publishes no golden vectors for this path
(SURVEY.md section 4), so the pins are outputs of the reference itself, generated
in the build container by ``oracle/gen_golden.py`` (which imports the unmodified
reference from /root/reference) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every oracle function against them, and
against the known-answer values listed in SURVEY.md appendix A.4.

Every function cites the reference file:line it restates (paths relative to the
reference root).
"""

from . import ops, loss, protos  # noqa: F401
