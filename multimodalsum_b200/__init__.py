"""mmsum-b200: B200-native (sm_100a) implementation of the MultimodalSum data-parallel training step."""
__version__ = "0.2.0"
