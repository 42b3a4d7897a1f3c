"""radmmm_b200 -- B200-native flow decoder + mel front end for RAD-MMM (drop-in for decoders.RADMMMFlow)."""
__version__ = "0.2.0"
