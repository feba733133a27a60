"""Input pipeline pieces that belong to the training hot path."""
from .packed import PackedBatchSpec, pack_batch, unpack_batch, mixlm_mask_sequence  # noqa: F401
