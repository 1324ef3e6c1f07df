"""tinydiff -- B200-native (sm_100a) DDPM hot path behind the call sites of
david-wb/tiny-diffusion.  Import as ``tinydiff`` (the directory name carries a hyphen):

    from tinydiff.diffusion import NoiseModel, ForwardProcess, sample
"""
__version__ = "0.1.0"
