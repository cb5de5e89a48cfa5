"""pram_b200: B200-native (sm_100a) implementation of PRAM's per-frame localization hot path."""
__version__ = '0.1.0'
