from vla_touch_b200.bridge_controller_no_visual import *  # noqa: F401,F403
from vla_touch_b200.bridge_controller_no_visual import DiffusionController, load_bridge_controller  # noqa: F401
