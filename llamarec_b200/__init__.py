"""llamarec_b200 -- B200-native (sm_100a) implementation of LlamaRec's stage-1 candidate-generation hot
path (LRURec encode -> catalogue score -> top-k -> metrics -> candidate emission) and the stage-2
verbalizer tail.  See DESIGN.md / INTEGRATION.md."""
from .model import LRURec, merge_lists  # noqa: F401
from .metrics import absolute_recall_mrr_ndcg_for_ks, absolute_metrics_batch_wrapper  # noqa: F401
from .retriever import LRURetriever  # noqa: F401
from .verbalizer import ManualVerbalizer  # noqa: F401
from . import stage2  # noqa: F401
from .evalset import DeviceEvalSet  # noqa: F401
from .sharded import CudaBackend, ShardedRetriever, shard_range  # noqa: F401

__all__ = ["LRURec", "merge_lists", "absolute_recall_mrr_ndcg_for_ks", "absolute_metrics_batch_wrapper",
           "LRURetriever", "ManualVerbalizer", "CudaBackend", "ShardedRetriever", "shard_range", "DeviceEvalSet"]
