from .attend import AttentionIntermediates
from .attention import Attention, AttentionConfig, AttentionSharedIntermediates
from .embeddings import (DiscreteContinuousEmbedding, DiscreteDenseContinuousEmbedding, AbsolutePositionalEmbedding,
                         ALiBiPositionalBias, LearnedALiBiPositionalBias)
from .feedforward import FeedForward, FeedForwardConfig, GLU
from .transformer import (TransformerConfig, EncoderConfig, DecoderConfig, TransformerRegistry, Transformer, Encoder, Decoder,
                          TransformerIntermediates)
