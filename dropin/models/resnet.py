from marconet_b200.models.resnet import *  # noqa: F401,F403
