"""daspeech_b200 -- B200-native (sm_100a) implementation of the DAG-loss hot path of ictnlp/DASpeech.

Public surface = the operator surface of the reference plugin's `DASpeech/custom_ops` package
(custom_ops/__init__.py:1): the same eight names, same argument meaning, same autograd contracts.

Beyond that surface (the callers either side of the path; imported explicitly):
`daspeech_b200.posterior` (alignment posterior of the S2S criterion), `daspeech_b200.glat` (GLAT force-emit masking of the
NAT criterion), `daspeech_b200.prefetch` (host -> device input prefetch), `daspeech_b200.dist` (utterance sharding).
"""
from .custom_ops import (  # noqa: F401
    dag_loss,
    dag_loss_with_alpha_beta,
    dag_best_alignment,
    dag_logsoftmax_gather_inplace,
    torch_dag_loss,
    torch_dag_best_alignment,
    torch_dag_logsoftmax_gather_inplace,
    logsumexp_keepdim,
)

__version__ = "0.1.0"
