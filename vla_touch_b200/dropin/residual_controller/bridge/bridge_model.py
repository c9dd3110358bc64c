from vla_touch_b200.bridge.bridge_model import StochasticInterpolants  # noqa: F401
from vla_touch_b200.ema import ExponentialMovingAverage  # noqa: F401
