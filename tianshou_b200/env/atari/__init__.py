from .atari_network import DQNet, ScaledObsInputActionReprNet, scale_obs

__all__ = ["DQNet", "ScaledObsInputActionReprNet", "scale_obs"]
