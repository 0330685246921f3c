"""rusty_marcher_b200 -- B200-native render hot path of rusty-marcher behind the engine crate's own
scene / geometry / render interface.  Module names follow engine/src/*.rs.

The compute path is the CUDA library librm_b200.so (csrc/, C ABI in include/rm_b200.h).  There is
no CPU fallback: rendering without an sm_100 GPU raises.
"""
from . import _abi, framebuffer, geometry, lights, obj, polygon, renderer, scene, shapes, sphere  # noqa: F401
from ._abi import RM_FP32, RM_FP64, RmError, init, shutdown  # noqa: F401
from .framebuffer import FrameBuffer, create_frame_buffer  # noqa: F401
from .geometry import Vec3f  # noqa: F401
from .lights import create_light  # noqa: F401
from .renderer import Renderer, create_renderer  # noqa: F401
from .scene import Scene  # noqa: F401
from .shapes import Reflectance  # noqa: F401
