"""audiolazy_b200 -- B200-native implementation of AudioLazy's linear-filter hot path."""
__version__ = "0.1.0"
