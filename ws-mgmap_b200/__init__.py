"""B200-native map update for WS-MGMap (drop-in for vlnce_baselines.common.rgb_mapping)."""
__version__ = "0.1.0"
