"""bnn_b200: B200-native (sm_100a) hot path of Alaya-in-Matrix/BNN behind the reference's own
Python interface.  Importing the package does not load the CUDA library; the first op does, and
raises if it is missing (there is no CPU fallback)."""
from .layout import Arch          # noqa: F401
from .bank import SampleBank      # noqa: F401

__all__ = ["Arch", "SampleBank", "BNN", "BNN_SGDMC", "BNN_BBB", "BNN_Dropout", "BNN_CDropout", "BNN_SVI", "ops"]
