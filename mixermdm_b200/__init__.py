"""mixermdm_b200 — B200-native (sm_100a) implementation of MixerMDM's batched denoising loop behind the
reference's Python model API. See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
