from .loss import (MSELossFunc, MSELoss_Landmark, CrossEntropyLoss2d, MSELoss, ResidualKDLoss)  # noqa: F401
