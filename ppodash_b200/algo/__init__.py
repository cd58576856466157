from .a2c_acktr import A2C_ACKTR  # noqa: F401
from .ppo import PPO  # noqa: F401
