"""radvlm_b200 — B200-native (sm_100a) implementation of RadVLM's multimodal encode path.

Only what the path needs: ``csrc/`` (CUDA kernels + C ABI), and the Python host mirror of the
reference's interface for this path:

  * :mod:`radvlm_b200.mm_arch`  — ``encode_images`` / ``prepare_inputs_labels_for_multimodal`` drop-ins
  * :mod:`radvlm_b200.mm_utils` — ``process_anyres_image`` / ``process_images`` on the GPU
  * :mod:`radvlm_b200.planner`  — grid selection, unpad / pool geometry, splice layout (bit-exact)
  * :mod:`radvlm_b200.encoder`  — weight packing + the SigLIP tower / projector runner
  * :mod:`radvlm_b200.dist`     — image sharding across ranks and the visual-token all-gather
"""
from . import _lib  # noqa: F401
from .planner import IGNORE_INDEX, IMAGE_TOKEN_INDEX  # noqa: F401

__all__ = ["_lib", "IGNORE_INDEX", "IMAGE_TOKEN_INDEX"]
__version__ = "0.1.0"
