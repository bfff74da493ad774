"""t2i_clip-gan_b200 -- B200-native DAMSM word/sentence matching loss (drop-in for the hot path of
dgjun32/T2I_CLIP-GAN: miscc/losses.py words_loss / sent_loss, GlobalAttention.func_attention).

The directory name is not a Python identifier; import it with

    import importlib; damsm = importlib.import_module("t2i_clip-gan_b200")

or use the identifier alias ``damsm_b200`` at the repository root.  To switch an existing reference checkout over,
call ``patch_reference()`` (patch.py): it replaces ``words_loss`` / ``sent_loss`` / ``func_attention`` inside the
reference's own ``miscc.losses`` / ``GlobalAttention`` modules and leaves everything else of those packages alone.
"""
from . import _lib
from ._lib import DamsmError
from .engine import CudaEngine, get_engine
from .patch import patch_reference, unpatch_reference
from .ops import (DEFAULT_GAMMAS, clip_resize, generator_regions, DamsmFuncAttention, DamsmNTXent, DamsmSentLoss, DamsmWordsLoss, LazyAttnMaps,
                  combine_column_lse, func_attention, nt_xent, project_regions, r_precision_scores, rm_special_token, sent_loss, standard_ntxent_mask,
                  words_loss)

__all__ = ["words_loss", "sent_loss", "func_attention", "DamsmWordsLoss", "DamsmSentLoss", "DamsmFuncAttention",
           "LazyAttnMaps", "nt_xent", "rm_special_token", "project_regions", "r_precision_scores", "DamsmNTXent", "standard_ntxent_mask", "CudaEngine", "get_engine", "DamsmError", "DEFAULT_GAMMAS", "combine_column_lse",
           "patch_reference", "unpatch_reference", "clip_resize", "generator_regions"]
__version__ = "0.1.0"
