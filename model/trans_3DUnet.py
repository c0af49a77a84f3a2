"""Drop-in for the reference's model/trans_3DUnet.py public names (Model_Dict, get_model_dict,
MaskTransUnet; reference lines 150-222), backed by lintransunet_b200."""
from lintransunet_b200.unet import MaskTransUnet, Model_Dict, get_model_dict  # noqa: F401
