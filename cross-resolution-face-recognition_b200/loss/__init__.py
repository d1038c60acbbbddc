from .loss import MSELossFunc, MSELoss_Landmark, CrossEntropyLoss2d  # noqa: F401
