"""t2i_clip-gan_b200 -- B200-native DAMSM word/sentence matching loss (drop-in for the hot path of
dgjun32/T2I_CLIP-GAN: miscc/losses.py words_loss / sent_loss, GlobalAttention.func_attention).

The directory name is not a Python identifier; import it with

    import importlib; damsm = importlib.import_module("t2i_clip-gan_b200")

or use the identifier alias ``damsm_b200`` at the repository root, or put this directory on ``sys.path`` and
keep the reference's own import lines (``from miscc.losses import sent_loss, words_loss``;
``from GlobalAttention import func_attention``).
"""
from . import _lib
from ._lib import DamsmError
from .engine import CudaEngine, get_engine
from .ops import (DEFAULT_GAMMAS, DamsmFuncAttention, DamsmNTXent, DamsmSentLoss, DamsmWordsLoss, LazyAttnMaps,
                  combine_column_lse, func_attention, nt_xent, project_regions, r_precision_scores, rm_special_token, sent_loss, standard_ntxent_mask,
                  words_loss)

__all__ = ["words_loss", "sent_loss", "func_attention", "DamsmWordsLoss", "DamsmSentLoss", "DamsmFuncAttention",
           "LazyAttnMaps", "nt_xent", "rm_special_token", "project_regions", "r_precision_scores", "DamsmNTXent", "standard_ntxent_mask", "CudaEngine", "get_engine", "DamsmError", "DEFAULT_GAMMAS", "combine_column_lse"]
__version__ = "0.1.0"
