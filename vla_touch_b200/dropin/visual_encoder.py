from vla_touch_b200.visual_encoder import DINOv2Encoder  # noqa: F401
