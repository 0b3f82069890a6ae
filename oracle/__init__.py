"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's training hot path (eladhoffer/convNet.pytorch: models/resnet.py,
trainer.py, utils/optim.py, utils/regularization.py, utils/cross_entropy.py) used as the parity checker.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package; nothing under ``convnet/`` does.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the reference itself, generated in the build container by ``oracle/make_golden.py``
(which imports /root/reference unmodified, behind import shims only) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks the restatement against those vectors.
"""
