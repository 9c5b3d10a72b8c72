"""Shapes of the reference's default DiTTO ("repo default model config").

Same attribute names as ``ConfigDiTTO`` in the reference (src/utils/Config.py:102-125) so that
``TrainDiTTO.py``-style scripts keep working; only the model/diffusion attributes the hot path reads."""


class ConfigDiTTO:
    MODEL_NAME = "DiTTO"
    HIDDEN_DIM = 768        # Config.py:109
    NUM_LAYERS = 5          # Config.py:110
    NUM_HEADS = 1           # Config.py:111
    TIME_DIM = 256          # Config.py:112
    TEXT_EMBED_DIM = 768    # Config.py:113
    DIFFUSION_STEPS = 1000  # Config.py:116
    MAX_TOKEN_LENGTH = 1024  # Config.py:124 (latent/text length cap of the data path)
    FRAME_RATE = 75         # EnCodec-24k: 24000 / 320 latent frames per second (EnCodec.py:33-37)
    DEVICE = "cuda"
