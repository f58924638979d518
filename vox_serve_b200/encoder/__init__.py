"""Audio-prompt encoders (``vox_serve/encoder``): the GLM-4-Voice speech tokenizer on the B200 kernels."""
from .glm import GLMEncoderConfig, GLMVoiceEncoder, GLMWhisperVQEncoder  # noqa: F401
