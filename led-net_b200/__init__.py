"""led-net_b200: B200-native LED-Net hot path behind the mmseg registry surface."""
__version__ = '0.1.0'
