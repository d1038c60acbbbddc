from .FSRnet import (Bottleneck, Coarse_SR_Network, Fine_SR_Decoder, Fine_SR_Encoder, Hourglass,  # noqa: F401
                     Prior_Estimation_Network, SRNetwork, _Residual_Block)
