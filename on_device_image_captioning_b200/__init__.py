"""xnv2_b200: B200-native (sm_100a) ExpansionNet v2 captioning inference path.

Public surface mirrors the reference (nighting0le01/On_Device_Image_Captioning):
``End_ExpansionNet_v2``, ``ExpansionNet_v2``, ``E2E_ExpansionNet_Captioner``, ``EsembleCaptioningModel``.
"""
from .config import XNConfig, swin_l_384, features_only, swin_tiny_test  # noqa: F401


def __getattr__(name):
    # lazy: importing the package must not require torch.cuda or the built library
    if name in ("End_ExpansionNet_v2", "ExpansionNet_v2", "E2E_ExpansionNet_Captioner", "EsembleCaptioningModel"):
        from . import models
        return getattr(models, name)
    if name in ("Engine", "EnginePair"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
