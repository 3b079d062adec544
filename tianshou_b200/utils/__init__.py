from .statistics import RunningMeanStd
from .torch_utils import policy_within_training_step, torch_device, torch_train_mode

__all__ = ["RunningMeanStd", "policy_within_training_step", "torch_device", "torch_train_mode"]
