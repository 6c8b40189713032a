"""madm_b200 — B200-native (sm_100a) drop-in for MADM's diffusion feature-extraction backbone."""
__version__ = "0.1.0"
