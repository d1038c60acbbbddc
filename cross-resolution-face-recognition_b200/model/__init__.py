from .FSRnet import (OverallNetwork, Course_SR_Network, Fine_SR_Encoder, Prior_Estimation_Network, Fine_SR_Decoder,
                     weights_init)  # noqa: F401
