"""omnihuman-1-hack_b200 -- B200-native (sm_100a) engine for the one hot path of
johndpope/OmniHuman-1-hack: the Wan2.1 DiT forward and the WanVAE decode.

The directory name carries a hyphen (it mirrors the reference repo's name), so import it through
the `b200dit` alias module at the repository root:

    import b200dit
    eng = b200dit.DitEngine.from_module(wan_model)      # or b200dit.install(wan_model)
"""
from ._lib import B200Error, LIB_PATH, MAX_ITEMS  # noqa: F401
from .engine import (DitEngine, VaeEngine, flash_attention, flash_attention_backward, kernel_launches, linear, profile_collect,  # noqa: F401
                     profile_enable)
from . import discriminator, flops, omni, parallel, pipelines, solvers, synthetic  # noqa: F401
from .omni import AudioProcessor  # noqa: F401
from .discriminator import AptDiscriminator  # noqa: F401
from .solvers import (FlowDPMSolverMultistepScheduler, FlowUniPCMultistepScheduler, get_sampling_sigmas,  # noqa: F401
                      retrieve_timesteps)
from .wan_shim import install, install_t2v, install_vae, uninstall  # noqa: F401
