"""pixelbox_b200 -- B200-native similarity search for PixelBox's semantic_hashes table.

The product is the C-ABI shared library built from pixelbox_b200/csrc (see include/pixelbox_b200.h);
this package is the thin Python host binding used by tests, the bench and the multi-GPU shard driver.
"""
__version__ = "0.1.0"
