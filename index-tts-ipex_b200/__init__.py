"""index-tts-ipex_b200: B200-native BigVGAN vocoder decode for IndexTTS (drop-in for
`indextts.BigVGAN.models.BigVGAN` as used by `indextts/infer.py`).

Import as `index_tts_ipex_b200` (the repo-root shim maps the hyphenated directory name)."""
from . import capi  # noqa: F401
from .models import BigVGAN, KAISER_TAPS  # noqa: F401
from .activation1d import Activation1d, FusedAntiAliasActivation, forward as anti_alias_activation_forward  # noqa: F401

from .mel import MelSpectrogramFeatures, SpeakerEmbeddingCache, melscale_fbanks_htk  # noqa: F401
from .sharding import decode_sharded, shard_bounds  # noqa: F401

Generator = BigVGAN   # infer.py:19 imports it under this alias
