"""B200-native SelectiveUNet (UNet_B + SelectiveNet heads) training / evaluation hot path.

Host side mirrors the reference's Python surface (``model.UNet_B``, ``selective_loss``,
``utils.compute_metric.Evaluator``, ``utils.net_utils``); every FLOP runs in hand-written
sm_100a kernels reached through the C ABI in ``include/sunet_b200.h``.
"""
__version__ = "0.1.0"
