"""rnamsm_b200 -- B200-native RNA-MSM MSA-transformer forward (emb + 120 tied-row attention maps).

Python host side of ``librnamsm_b200.so`` (C ABI in ``include/rnamsm_b200.h``); mirrors the
reference's ``MSATransformer`` / ``AxialTransformerLayer`` API.  Importing this package loads
the CUDA library and raises if it is missing -- there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (loads librnamsm_b200.so, fails loudly when absent)
from .alphabet import Alphabet, Vocab, read_msa, tokenize_msa  # noqa: F401
from .model import MSATransformer  # noqa: F401
from .modules import (AxialTransformerLayer, ColumnSelfAttention, ContactPredictionHead,  # noqa: F401
                      FeedForwardNetwork, LearnedPositionalEmbedding, NormalizedResidualBlock, RobertaLMHead,
                      RowSelfAttention)
from .inference import (extract_features, extract_features_batch, extract_features_batch_streamed,  # noqa: F401
                        extract_features_streamed, pack_rsa_input, pack_ss_input, plan_batches, run_inference)
from .ingest import ingest_msa  # noqa: F401

__all__ = ["MSATransformer", "AxialTransformerLayer", "RowSelfAttention", "ColumnSelfAttention",
           "FeedForwardNetwork", "NormalizedResidualBlock", "LearnedPositionalEmbedding", "RobertaLMHead",
           "ContactPredictionHead", "Alphabet", "Vocab", "read_msa", "tokenize_msa", "extract_features",
           "run_inference", "ingest_msa", "pack_ss_input", "extract_features_streamed", "extract_features_batch",
           "extract_features_batch_streamed", "pack_rsa_input", "plan_batches"]
