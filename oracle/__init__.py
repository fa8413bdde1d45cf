"""TEST INFRASTRUCTURE: CPU restatement (PyTorch fp32 + numpy float64) of the reference's
per-step part-disentanglement path.  Importable only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; the product never routes through it.

Pinning: the reference (TensorFlow 1.14 graph code) cannot be imported here and its own
tests hold no vectors for this path.  tests/golden/make_golden.py executes the reference's
own function bodies (read from /root/reference at generation time, never copied) under a
TF1->torch eager shim and commits the outputs as fixtures; tests/test_oracle_golden.py pins
this package against them.  TF's kernel-level rounding (Eigen exp/log, cuBLAS matmul order,
fp32 matrix_inverse) is emulated by torch CPU ops in those fixtures, so the pin is on the
algorithm, not on TF's last ulp.
"""
from . import canon, parts, tps, step, stats, priors, ingest, inject_conv  # noqa: F401
