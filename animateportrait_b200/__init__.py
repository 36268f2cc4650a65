"""B200-native (sm_100a) implementation of AnimatePortrait's Module2 generator hot path.

Public surface (mirrors Module2/models/networks.py for this one generator):
    define_G, ResnetConditionTriGenerator32_full_ifw, get_norm_layer, init_weights, install
plus `frames.render_frames` / `frames.render_frames_sharded` for clips sharded over GPUs,
`compose.blend_and_convert` for the blend + uint8 output stage that follows the generator,
`conditioning.{draw2, cal_motion256, kp_to_map_some, matte_photo}` for the per-frame inputs the reference makes on the
CPU (Module2/data/umlvdfw_test_dataset.py, Module2/models/geomcgt_ifw_test_model.py) and `clip.ClipRenderer` /
`clip.render_clip_sharded`, which string all of it together for one photo and a landmark track.
"""
from .netg import (NETG_NAME, ResnetBlock, ResnetBlock2, ResnetConditionTriGenerator32_full_ifw, conv2d_debug,
                   define_G, get_norm_layer, init_weights, install)

__all__ = ["NETG_NAME", "ResnetBlock", "ResnetBlock2", "ResnetConditionTriGenerator32_full_ifw", "conv2d_debug",
           "define_G", "get_norm_layer", "init_weights", "install"]
