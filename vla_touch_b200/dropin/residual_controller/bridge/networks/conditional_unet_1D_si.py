from vla_touch_b200.bridge.networks.conditional_unet_1D_si import (DiffusionConditionalUnet1D,  # noqa: F401
                                                                      InterpolantsConditionalUnet1D)
