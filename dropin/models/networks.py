from marconet_b200.models.networks import *  # noqa: F401,F403
