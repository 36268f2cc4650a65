"""B200-native (sm_100a) implementation of AnimatePortrait's Module2 generator hot path.

Public surface (mirrors Module2/models/networks.py for this one generator):
    define_G, ResnetConditionTriGenerator32_full_ifw, get_norm_layer, init_weights, install
plus `frames.render_frames` / `frames.render_frames_sharded` for clips sharded over GPUs and
`compose.blend_and_convert` for the blend + uint8 output stage that follows the generator.
"""
from .netg import (NETG_NAME, ResnetBlock, ResnetBlock2, ResnetConditionTriGenerator32_full_ifw, conv2d_debug,
                   define_G, get_norm_layer, init_weights, install)

__all__ = ["NETG_NAME", "ResnetBlock", "ResnetBlock2", "ResnetConditionTriGenerator32_full_ifw", "conv2d_debug",
           "define_G", "get_norm_layer", "init_weights", "install"]
