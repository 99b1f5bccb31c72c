"""nerficg_b200 -- B200-native (sm_100a) drop-in for nerficg's vanilla-NeRF volume-rendering hot path."""
__version__ = '0.1.0'
