from marconet_b200.models.textvit_arch import *  # noqa: F401,F403
