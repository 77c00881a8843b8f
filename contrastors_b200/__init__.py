"""contrastors_b200: the contrastive-training hot path of nomic-ai/contrastors, rebuilt for B200 (sm_100a).

Python surface = the reference's names (``clip_loss``, ``cache_loss``, ``grad_cache_loss``, ``gather_with_grad``,
``LogitScale``, ``BiEncoder``, ``DualEncoder``); the work underneath is hand-written CUDA behind a C ABI
(``include/contrastors_b200.h``, built into ``libcontrastors_b200.so``).  No CPU fallback, no Triton, no dispatch.
"""
from . import _lib  # noqa: F401
from .distributed import gather, gather_with_grad  # noqa: F401
from .logit_scale import LogitScale  # noqa: F401
from .loss import (accumulate_gradients, cache_loss, clip_loss, get_chunked_embeddings, grad_cache_loss,  # noqa: F401
                   matryoshka_clip_loss, symmetric_clip_loss)
from .models import BiEncoder, BiEncoderConfig, NomicBertConfig, NomicBertModel, nomic_bert_base  # noqa: F401
from .rand_state import RandContext  # noqa: F401
from . import checkpoint, trainer  # noqa: F401
from .vit import DualEncoder, ViTConfig, ViTModel, VisionBiEncoder, VisionBiEncoderConfig, vit_b16, vit_l14  # noqa: F401

__version__ = "0.1.0"
