"""B200-native (sm_100a) implementation of Grasp-Any-Region's region-captioning inference path.

Public surface = the reference's: ``GARModel`` / ``GARConfig`` (hf_models/modeling_gar.py,
configuration_gar.py).  Importing this package never touches CUDA; using it without the built
library or without a GPU raises (no CPU / PyTorch fallback)."""
import os as _os

# Pipelined serving (GARModel.generate_batches) runs a decode stream next to an encode stream; they must not share a
# hardware work queue (the default is 8 queues for all streams of the process).  Only effective before CUDA initialises.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .config import GARSpec, gar_1b, gar_8b, gar_8b_mini, gar_small, gar_tiny  # noqa: F401

__all__ = ["GARModel", "GARConfig", "GARSpec", "gar_1b", "gar_8b", "gar_tiny", "gar_small"]


def __getattr__(name):
    if name == "GARModel":
        from .modeling_gar import GARModel

        return GARModel
    if name == "GARConfig":
        from .configuration_gar import GARConfig

        return GARConfig
    raise AttributeError(name)
