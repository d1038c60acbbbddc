from .eval import accuracy  # noqa: F401
from .utils import calculate_accuracy, calculate_roc, cosine_identify, l2_norm  # noqa: F401
