from .constructor import Constructor, ModuleConfig, VariableModuleConfig, Registry
from .layers import Residual, AdaptiveLayerNorm, LayerNorm
from .sampling import top_k, top_p, top_a, filter_logits_and_sample
from .transformer import *  # noqa: F401,F403
