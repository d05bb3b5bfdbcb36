from marconet_b200.models.ocr import *  # noqa: F401,F403
