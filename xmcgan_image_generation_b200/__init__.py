"""B200-native (sm_100a) implementation of XMC-GAN's data-parallel train_step hot path behind the reference's
Python call surface. See DESIGN.md / INTEGRATION.md."""
__all__ = ["configs", "engine", "libml", "nets", "ops", "parallel", "train_utils", "xmc_gan"]
