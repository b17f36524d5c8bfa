from .config import ModelConfig, PRESETS, get_config
