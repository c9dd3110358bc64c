from vla_touch_b200.bridge_controller import *  # noqa: F401,F403
from vla_touch_b200.bridge_controller import DiffusionController, load_bridge_controller  # noqa: F401
