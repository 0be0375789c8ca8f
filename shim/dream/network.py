"""dream.network -> dream_b200.network (same names, dream/network.py:18-696)."""
from dream_b200.network import *          # noqa: F401,F403
from dream_b200.network import (KNOWN_ARCHITECTURES, KNOWN_OPTIMIZERS, DreamNetwork,   # noqa: F401
                                create_network_from_config_data, create_network_from_config_file)
