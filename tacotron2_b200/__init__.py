"""tacotron2_b200 -- a B200-native (sm_100a) Tacotron 2 mel-spectrogram engine behind the
NVIDIA/tacotron2 nn.Module API.  See DESIGN.md / INTEGRATION.md."""
from ._engine import dropout_masks, invalidate_weights  # noqa: F401
from .hparams import create_hparams  # noqa: F401
from .layers import TacotronSTFT  # noqa: F401
from .loss_function import Tacotron2Loss  # noqa: F401
from .model import Decoder, Encoder, Postnet, Tacotron2  # noqa: F401
from . import amp  # noqa: F401
from .optim import AmpFusedClipAdam, FusedClipAdam  # noqa: F401

__all__ = ["Tacotron2", "Encoder", "Decoder", "Postnet", "Tacotron2Loss", "create_hparams", "dropout_masks",
           "FusedClipAdam", "AmpFusedClipAdam", "amp", "invalidate_weights", "TacotronSTFT"]
