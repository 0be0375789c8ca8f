"""dream_b200 -- B200-native implementation of the DREAM belief-map hot path."""
__version__ = "0.1.0"
