"""Import-path shim: the reference scripts do `from model.trans_3DUnet import get_model_dict`
(train3D.py:20, inference_embed_attn.py:14).  Putting this repository on sys.path ahead of the
reference makes those scripts pick up the B200 implementation without edits."""
